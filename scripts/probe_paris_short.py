import sys, os
sys.argv = [sys.argv[0], "none"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_n2.py")).read())
run("garch", "paris", 16384, 64, 60, hetero=True)
run("svm", "paris", 16384, 64, 60, hetero=True)
run("lgssm", "paris", 16384, 64, 60, hetero=True)
run("garch", "paris", 16384, 1, 60)
