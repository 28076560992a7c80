import os, sys, runpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ismcsolver_b200.build as b
b.LIB_PATH = os.path.abspath(sys.argv[1]); b.is_stale = lambda: False
b.build_library = lambda *a, **k: b.LIB_PATH
sys.argv = ["bench.py"] + sys.argv[2:]
runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), run_name="__main__")
