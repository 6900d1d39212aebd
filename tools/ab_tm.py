import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import quick_bench as q
for tm in (1,0,1,0):
    q.run(1080,1920,(64,64),(32,32),101,tmem=tm)
q.run(2160,3840,(64,64),(32,32),41,tmem=1)
q.run(2160,3840,(64,64),(32,32),41,tmem=0)
q.run(1080,1920,(64,64),(48,48),41,tmem=1)
q.run(1080,1920,(64,64),(48,48),41,tmem=0)
