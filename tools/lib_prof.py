"""dev aid: tools/quick_prof.py against an alternative build of the library (MPPI_LIB=path)"""
import os, sys, runpy
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from mppi_gpu_b200 import capi
if os.environ.get("MPPI_LIB"):
    capi.LIB_PATH = os.path.abspath(os.environ["MPPI_LIB"])
sys.argv = ["quick_prof.py"] + sys.argv[1:]
runpy.run_path(os.path.join(os.path.dirname(__file__), "quick_prof.py"), run_name="__main__")
