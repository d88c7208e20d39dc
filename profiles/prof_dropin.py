"""cProfile of one drop-in handler step (host side): python profiles/prof_dropin.py [precision]"""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda", 0)
feats = [torch.randn(16384, 1024, device=dev) for _ in range(16)]
print(bench.handler_leg("advmil_b200", dev, feats, 2, 2, precision=prec))
pr = cProfile.Profile()
pr.enable()
r = bench.handler_leg("advmil_b200", dev, feats, 2, 1, precision=prec)
pr.disable()
print(r)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(60)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print(s.getvalue()[:5000])
