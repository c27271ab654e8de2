"""cfg4 timing: 512^3 density lattice of one synthetic frame on one GPU, with the per-entry-point kernel times."""
import torch, time, sys
sys.path.insert(0, "tests")
from util import synthetic
from anim_nerf_b200.anim_nerf import AnimNeRF
from anim_nerf_b200 import inference
net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=True, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0)).cuda()
posed_np, tmpl_np = synthetic.make_body_params(1, seed=5)
posed = {k: torch.from_numpy(v).cuda() for k, v in posed_np.items()}; tmpl = {k: torch.from_numpy(v).cuda() for k, v in tmpl_np.items()}
with torch.no_grad():
    net.setup_frame(posed, tmpl, None)
    out = torch.empty(512,512,512, device="cuda")
    for _ in range(2): inference.query_density_grid(net, 512, out=out)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(3): inference.query_density_grid(net, 512, out=out)
    torch.cuda.synchronize(); print("grid_512 ms", (time.time()-t)/3*1e3)
    from anim_nerf_b200 import _lib
    import numpy as np
    t = _lib.enable_timing(True)
    inference.query_density_grid(net, 512, out=out)
    torch.cuda.synchronize()
    for k, v in t.items():
        print("%-28s %3d calls %8.3f ms total" % (k, len(v), sum(a.elapsed_time(b) for a, b in v)))
    print("valid fraction", float((out > 0).float().mean()))
