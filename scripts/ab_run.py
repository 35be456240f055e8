"""Run a script against a variant build of the library: python scripts/ab_run.py <libfdcm_b200.so> <script.py> [args...]"""
import runpy
import sys
sys.path.insert(0, ".")
import openfdcm_b200._lib as _lib
_lib.LIB_PATH = sys.argv[1]
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
