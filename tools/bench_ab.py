"""A/B of bench.py against a variant build of the library on ONE box (box-to-box clocks differ by 10 % under the power cap):
    python tools/bench_ab.py --lib tools/variants/lib_old.so -- --steps 4 --no-cpu-baseline --no-eval"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--lib" in sys.argv:
    import adapter4rec_b200.lib as _lib
    i = sys.argv.index("--lib")
    _lib.LIB_PATH = os.path.abspath(sys.argv[i + 1])
    del sys.argv[i:i + 2]
if "--" in sys.argv:
    sys.argv.remove("--")
import bench
bench.main()
