"""Build art_tts_b200/lib/libmas_dbg.so: the product sources with -DMAS_DEBUG_CHECKS (MAS_CHECK assertions)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from art_tts_b200 import build as b

print(b.build(extra=["-DMAS_DEBUG_CHECKS"], out=os.path.join(b.LIBDIR, "libmas_dbg.so")))
