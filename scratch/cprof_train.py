import cProfile, pstats, sys, os, io
sys.argv = ["prof_train.py", "c3", "6"]
sys.path.insert(0, "/root/repo/profiles")
exec(open("/root/repo/profiles/prof_train.py").read())
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(60)
print(s.getvalue())
