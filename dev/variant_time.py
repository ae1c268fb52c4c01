"""dev/variant_time.py -- stage times of the C4 scene (full image and a 128-row band) for every library variant built by
dev/build_variants.sh (one subprocess per variant: XYZ_B200_LIB selects the library)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in sorted(glob.glob(os.path.join(ROOT, "xyz-autodiff-cuda_b200", "lib_variants", "*"))):
    so = os.path.join(d, "libxyz_b200.so")
    if not os.path.exists(so):
        continue
    env = dict(os.environ, XYZ_B200_LIB=so)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "dev", "fwd_threads_sweep.py")], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True).stdout
    for line in out.splitlines():
        if line.startswith("FWDSWEEP") and ("rows= 1024" in line or "rows=  128" in line):
            print(f"VARIANT {os.path.basename(d):18s} {line[len('FWDSWEEP threads=auto '):]}", flush=True)
