"""Build tuning variants of the CUDA library (other launch bounds) next to the product build: lib/var_<name>.so."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from allscale_ipic3d_b200 import build
for spec in sys.argv[1:]:
    name, defs = spec.split(":")
    print(build.build(force=True, defines=[d for d in defs.split(",") if d], name="var_%s.so" % name))
