"""A/B of two builds of the library on ONE box (boxes of the pool differ by a few percent): alternating subprocess runs of the
headline geometries.  usage: python tools/ab_lib.py libA.so libB.so"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import quick_bench as q; "
        "q.run(1080,1920,(64,64),(32,32),101); q.run(2160,3840,(64,64),(32,32),41); q.run(4320,7680,(128,128),(64,64),21)") % (os.path.dirname(here), here)
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, B2PIV_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout
        print("==", lib)
        print("\n".join(l[20:150] for l in out.splitlines() if "Mwin/s" in l), flush=True)
