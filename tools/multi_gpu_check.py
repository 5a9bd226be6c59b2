"""torchrun --nproc-per-node N tools/multi_gpu_check.py : the region and surface drivers sharded over N ranks
(NCCL) must reproduce the single-rank goldens bit for bit (1/2/4/8-GPU maps are identical arrays)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from rg_b200 import RG, make_param_dict
from rg_b200.rg_modules import region_maps, compute_cr_batch, shard_indices, gather_sharded

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ok = True
for name, L, ct in (("region_L5J5_12x12_expm_onestep", 5, "expm_onestep"), ("region_L10J10_sub16_expm_onestep", 10, "expm_onestep"),
                    ("region_L5J5_12x12_expm_multiply", 5, "expm_multiply")):
    g = np.load(os.path.join(G, name + ".npz"))
    case = RG(make_param_dict(L=L, J=L, CanonicalTransformation=ct, SaveData=False), device=local)
    data, info = region_maps(case, g["xs"], g["ys"])
    good = np.array_equal(data[:, :, 0], g["conv"]) and np.array_equal(data[:, :, 1], g["count"]) and np.array_equal(info, g["error"])
    ok &= good
    if rank == 0:
        print("%s on %d ranks: %s" % (name, world, "bit-exact" if good else "MISMATCH"))
    case.close()
# surface: columns sharded, float64 gather
g = np.load(os.path.join(G, "surface_L5J5_n8.npz"))
case = RG(make_param_dict(L=5, J=5, CanonicalTransformation="expm_onestep", RelDist=1e-5, Nxy=8, SaveData=False), device=local)
eps = np.linspace(0, 1, 8)
idx = shard_indices(8, rank, world)
loc = np.array(compute_cr_batch([eps[i] for i in idx], case), dtype=np.float64).reshape(len(idx), -1)
full = gather_sharded(loc.view(np.int64).reshape(len(idx), -1), 8, rank, world).view(np.float64)
good = np.allclose(full[:, 1], g["data"][1], rtol=2e-5, atol=0)
ok &= good
if rank == 0:
    print("surface_L5J5_n8 on %d ranks: %s" % (world, "ok" if good else "MISMATCH"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
