"""The reference's RegionList known-answer tests (test/alltests.cpp:38-114, test2.c2: four Join cases with their sizes)
against the oracle's restatement, plus IsOverlapped's boundary behaviour."""
import ctypes as C

import numpy as np

import fx


def _list(lib, regs):
    buf = np.zeros(2 * (len(regs) + 8), np.int32)
    n = 0
    for s, e in regs:
        n = lib.orc_regions_add(buf.ctypes.data_as(C.c_void_p), n, s, e)
    return buf, n


def _join(a, b):
    lib = fx.build_oracle()
    ba, na = _list(lib, a)
    bb, nb = _list(lib, b)
    ln = C.c_longlong(0)
    big = np.zeros(2 * (na + nb) + 16, np.int32)
    big[: 2 * na] = ba[: 2 * na]
    n = lib.orc_regions_join(big.ctypes.data_as(C.c_void_p), na, bb.ctypes.data_as(C.c_void_p), nb, C.byref(ln))
    return [(int(big[2 * i]), int(big[2 * i + 1])) for i in range(n)], int(ln.value)


def test_reference_kats_region_join():
    assert _join([(100, 200)], [(150, 160)]) == ([(150, 160)], 11)
    assert _join([(300, 400)], [(250, 350)]) == ([(300, 350)], 51)
    assert _join([(100, 155), (155, 200), (300, 400), (500, 600)],
                 [(150, 160), (250, 350), (550, 650), (650, 750)]) == ([(150, 160), (300, 350), (550, 600)], 113)
    assert _join([(990130, 990630), (1020346, 1022346)], [(989819, 989939), (990162, 990402)]) == ([(990162, 990402)], 241)


def test_region_overlap_boundaries():
    lib = fx.build_oracle()
    buf, n = _list(lib, [(10, 20), (15, 30), (40, 50)])
    ln = C.c_longlong(0)
    n = lib.orc_regions_collapse(buf.ctypes.data_as(C.c_void_p), n, C.byref(ln))
    assert [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)] == [(10, 30), (40, 50)] and ln.value == 32
    inside = [p for p in range(0, 60) if lib.orc_regions_overlapped(buf.ctypes.data_as(C.c_void_p), n, p)]
    assert inside == list(range(10, 31)) + list(range(40, 51))
