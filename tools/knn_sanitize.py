import sys; sys.path.insert(0, '.')
import numpy as np, torch
from cofii2p_b200 import ops
from oracle import knn as ok
rng = np.random.default_rng(0)
src = (rng.normal(size=(700,3))*10).astype(np.float32); qry = (rng.normal(size=(300,3))*10).astype(np.float32)
for mode in (0,1):
    got = ops.knn_table(torch.from_numpy(src).cuda(), torch.from_numpy(qry).cuda(), 1, 128, mode).cpu().numpy()
    print(mode, np.array_equal(got, ok.knn_table(src, qry, 128, mode)))
lv = [torch.from_numpy(src).cuda(), torch.from_numpy(src[:350]).cuda()]
o = ops.knn_pyramid(lv, 1, 128, 0); torch.cuda.synchronize(); print('pyr ok')
