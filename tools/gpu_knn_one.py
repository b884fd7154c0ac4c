"""PatchCore-sized nearest-neighbour search (784 x 16384 x 1536) a few times, for ncu captures and timings (development aid)."""
import sys, time, torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import producers
g = torch.Generator().manual_seed(9)
x, bank = torch.randn(784, 1536, generator=g).cuda(), torch.randn(16384, 1536, generator=g).cuda()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc, loc = producers.knn_min(x, bank)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"knn 784 x 16384 x 1536: {dt*1e3:.3f} ms wall (prep + search + finish), {3*2*784*16384*1536/dt/1e12:.1f} bf16 TFLOP/s incl. the hi/lo split")
