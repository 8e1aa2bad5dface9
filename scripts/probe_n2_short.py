import sys, os
sys.argv = [sys.argv[0], "none"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_n2.py")).read())
run("svm", "poyiadjis_N2", 16384, 8, 10)
run("svm", "poyiadjis_N2", 65536, 4, 6)
run("garch", "poyiadjis_N2", 16384, 8, 10)
