"""Short throughput probe: SVM poyiadjis_N f32 sorted, N=65536, B in {128, 512}."""
import sys, os
sys.argv = [sys.argv[0]]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "quick_probe.py")).read().split("for N, B in")[0])
for N, B in [(65536, 128), (65536, 512), (65536, 512)]:
    run("svm", "poyiadjis_N", N, B, "f32", "multinomial_sorted", reps=5)
