import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
for th in (16, 32, 64, 128):
    rps, t, dt = bench.cpu_reference_rays_per_s(256, repeats=1, threads=th)
    print(f"threads {th}: {rps:.2f} rays/s ({dt:.1f} s for 256 rays)", flush=True)
