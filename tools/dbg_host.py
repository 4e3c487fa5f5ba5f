import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppo_radiotherapy_b200 as rt
DEV = "cuda:0"
n = 257
tids = (np.arange(n) * 7919 % 1000).astype(np.int32)[None, :]
rng = np.random.default_rng(2)
acts = rng.uniform(-1, 1, (6, n, 6)).astype(np.float32)
def cu(a): return torch.as_tensor(np.ascontiguousarray(a), device=DEV)
for variant in ("plain", "side", "side_big"):
    a = rt.BatchedEpisodes(n, device=DEV); a.set_tumour_schedule(tids)
    b = rt.BatchedEpisodes(n, device=DEV); b.set_tumour_schedule(tids)
    h_obs, h_rew = np.empty((n, 9), np.float32), np.empty(n, np.float64)
    h_term, h_trunc = np.empty(n, np.uint8), np.empty(n, np.uint8)
    big = torch.empty(1 << 26, dtype=torch.float32, device=DEV)
    side = torch.cuda.Stream(DEV)
    a.reset(); a.step(cu(acts[0]), want_info=False)
    if variant == "plain":
        b.reset(); b.step(cu(acts[0]), want_info=False)
    else:
        with torch.cuda.stream(side):
            if variant == "side_big":
                for _ in range(20): big.add_(1.0)
            b.reset(); b.step(cu(acts[0]), want_info=False)
    torch.cuda.synchronize()
    print(variant, "after step0 pose equal:", torch.equal(a.pose(), b.pose()), "obs equal:", torch.equal(a.obs, b.obs))
    a.step(cu(acts[1]), want_info=False)
    b.step_host(acts[1], h_obs, h_rew, h_term, h_trunc)
    print(variant, "after host step obs equal:", np.array_equal(h_obs, a.obs.cpu().numpy()), "pose equal:", torch.equal(a.pose(), b.pose()))
    a.close(); b.close()
