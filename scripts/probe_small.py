import sys, os
sys.argv = [sys.argv[0]]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "quick_probe.py")).read().split("for N, B in")[0])
for N, B in [(1024, 1), (1024, 64), (1024, 1024), (2048, 1024)]:
    run("svm", "poyiadjis_N", N, B, "f32", "multinomial_sorted", reps=5)
