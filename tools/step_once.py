"""Runs bench.py's step a few times (for ncu launch lists): python tools/step_once.py [steps]"""
import sys
sys.argv = ["bench.py", "--steps", sys.argv[1] if len(sys.argv) > 1 else "2", "--warmup", "3", "--no-cpu-baseline", "--no-graph", "--no-f32-line", "--repeats", "1"] + sys.argv[2:]
sys.path.insert(0, ".")
import runpy
runpy.run_path("bench.py", run_name="__main__")
