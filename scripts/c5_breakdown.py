"""Diagnostic (torchrun): split one slab-decomposed ETDRK2 step into compute-only and transpose-only time."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import exponax_b200 as ex
from exponax_b200 import _slab

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
_slab.init_process_group_nccl(torch.cuda.current_device())
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
slab = ex.SlabStepper.navier_stokes_velocity(2 * np.pi, N, 1e-3, injection_mode=4)
slab.plan()
n = N // world
uh = torch.randn((3, N, n, N // 2 + 1, 2), device="cuda").mul_(1e-3)
uh = torch.view_as_complex(uh).contiguous()

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return 1e3 * (time.perf_counter() - t0) / reps

res = {}
for ov in (False, True):
    slab.overlap = ov
    res[f"full_overlap={ov}"] = timeit(lambda: slab.step_fourier(uh, inplace=True))
# transposes only
a = torch.zeros((1, n, N, N // 2 + 1), dtype=torch.complex64, device="cuda")
b = torch.zeros((1, N, n, N // 2 + 1), dtype=torch.complex64, device="cuda")
def only_transposes():
    for _ in range(2):
        for _ in range(6):
            _slab.transpose_b_to_a(b, out=a)
        for _ in range(3):
            _slab.transpose_a_to_b(a, out=b)
res["transposes_only"] = timeit(only_transposes)
def only_a2a():
    send = b[0].view(world, n, n, N // 2 + 1)
    recv = torch.empty_like(send)
    for _ in range(18):
        dist.all_to_all_single(recv, send)
res["all_to_all_only"] = timeit(only_a2a)
# compute only: no-op transposes
orig = (_slab.transpose_a_to_b, _slab.transpose_b_to_a)
_slab.transpose_a_to_b = lambda x, group=None, out=None: out
_slab.transpose_b_to_a = lambda x, group=None, out=None: out
slab.overlap = False
res["compute_only"] = timeit(lambda: slab.step_fourier(uh, inplace=True))
if rank == 0:
    print(json.dumps({k: round(v, 2) for k, v in res.items()}))
dist.destroy_process_group()
