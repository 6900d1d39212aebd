import sys, os
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo")
import quick_bench as q
which = sys.argv[1]
if which == "pad": q.run(1080, 1920, (26, 26), (12, 12), 11, reps=2, variant=4)
elif which == "pad101": q.run(1080, 1920, (26, 26), (12, 12), 101, reps=3, variant=4)
elif which == "nat11": q.run(1080, 1920, (64, 64), (32, 32), 11, reps=3, variant=2)
elif which == "nat11u": q.run(1080, 1920, (64, 64), (40, 40), 11, reps=3, variant=2)
