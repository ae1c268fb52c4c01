# dev/splat_knob.sh -- C4 iteration time over the CTAs-per-SM knob of the counting-sort binning
for k in 2 3 4 6 8 12; do XYZ_SPLAT_BIN_CTAS_PER_SM=$k python dev/splat_time.py "ctas/sm=$k" | grep "C4 splat"; done
