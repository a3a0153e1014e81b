"""``import spp`` -> the product package (whose directory name is not a Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("person-recognition-for-pose-estimation_b200")
