# dev/splat_knob3m.sh -- large-N behaviour of the counting-sort binning over its CTAs-per-SM knob
for k in 1 2 4; do echo "ctas/sm=$k"; XYZ_SPLAT_BIN_CTAS_PER_SM=$k python dev/splat_bin_crossover.py 800000 1600000 3000000; done
