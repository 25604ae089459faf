"""BW_L2 on this box: random 32/64/128-byte reads over L2-resident buffers, all SMs, best of 10.
Prints one JSON object; the bench's roofline uses the 8 MiB / 32-byte and 64-byte figures (fqb_measure_l2)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastquick_b200 import _abi  # noqa: E402

lib = _abi.load_library()
out = {"method": "fqb_measure_l2: 148x8 blocks x 256 threads, 512 independent ld.global.cg 256-bit loads per thread at LCG-random "
                 "unit-aligned offsets, CUDA events, 2 warm-up + 10 timed launches", "results": []}
for mib in (4, 8, 16, 64):
    for unit in (32, 64, 128):
        best, med = C.c_double(0), C.c_double(0)
        rc = lib.fqb_measure_l2(0, C.c_int64(mib << 20), unit, 10, C.byref(best), C.byref(med))
        assert rc == 0, lib.fqb_last_error()
        out["results"].append({"buffer_mib": mib, "unit_bytes": unit, "gbs_best": round(best.value, 1), "gbs_median": round(med.value, 1)})
print(json.dumps(out, indent=1))
