import sys, time
sys.path.insert(0, '/root/repo')
from emme_b200 import Input, pic
inp = Input('/root/repo/tests/golden/inputs/pic.json')
for k in range(3):
    t0 = time.perf_counter()
    r = pic.solve_once_pic(inp, seed=1)
    print('solve_once_pic', round(time.perf_counter() - t0, 4), r['timing'], flush=True)
