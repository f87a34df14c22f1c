"""Build a development variant of the library: tools/build_variant.py NAME DEFINE...  ->  pyds_b200/libpyds_b200_NAME.so
(load it with PYDSB_LIB=<path>; experiments only, the product is pyds_b200/libpyds_b200.so)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyds_b200 import build  # noqa: E402

name, defines = sys.argv[1], sys.argv[2:]
out = os.path.join(build.HERE, f"libpyds_b200_{name}.so")
print(build.build(force=True, defines=defines, out=out))
