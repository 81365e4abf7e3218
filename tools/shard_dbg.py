import os; os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import sys, threading, time
sys.path[:0]=['box2d-mt_b200/python','oracle','tests']
import numpy as np, b2cuda, b2cuda_types as T, b2shard, parity, ref, scenes
import test_sharding as ts
scene = scenes.pile(36, 8); scene.world_flags &= ~T.WORLD_CONTINUOUS
worlds, plans = ts._shard_worlds(b2cuda, scene, 2, margin=2.5)
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t=time.time()
    res=[None,None]
    def run(i):
        try: res[i]=worlds[i].step(pos_iters=1)
        except Exception as e: res[i]=e
    th=[threading.Thread(target=run,args=(i,)) for i in range(2)]
    [x.start() for x in th]; [x.join() for x in th]
    if any(isinstance(r, Exception) for r in res) or step % 25 == 0: print(step, round(time.time()-t,3), [ (r if isinstance(r,Exception) else (int(r['constraintCount']), int(r['kernelLaunches']))) for r in res], flush=True)
