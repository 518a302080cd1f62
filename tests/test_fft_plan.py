"""The FFT planner (csrc/fft.cuh: choose_radices_plan, through ax3d_fft_plan; host only): for every lucky number the
reference can ask for (PreloopFFTW.cpp:59-109) the radices multiply to Nr, come from the set the kernels implement
(2 ... 13 and 16), use the fewest stages any factorisation into such radices allows, and end with an odd radix whenever Nr has an odd factor; other sizes are refused."""
import ctypes as C
import functools

import pytest

from axisem3d_b200 import capi
from axisem3d_b200 import spectral as SP


RADICES = (2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16)


def plan(n):
    lib = capi.load(build_if_missing=False)
    r = (C.c_int * 10)()
    ns = C.c_int(0)
    rc = lib.ax3d_fft_plan(n, r, 10, C.byref(ns))
    if rc:
        raise RuntimeError(lib.ax3d_last_error().decode())
    return [r[k] for k in range(ns.value)]


@functools.lru_cache(None)
def min_stages(n):
    if n == 1:
        return 0
    best = 99
    for r in RADICES:
        if n % r == 0:
            best = min(best, 1 + min_stages(n // r))
    return best


def test_plans_of_all_lucky_numbers_up_to_2100():
    lucky = sorted({SP.next_lucky_number(n) for n in range(3, 2101)} | {SP.next_lucky_number(n, True) for n in range(3, 2101)})
    assert 208 in lucky and 416 in lucky and 2016 in lucky and 2025 in lucky
    for n in lucky:
        r = plan(n)
        prod = 1
        for x in r:
            assert x in RADICES
            prod *= x
        assert prod == n, (n, r)
        assert len(r) == min_stages(n), (n, r)
        if any(x % 2 for x in r):
            assert r[-1] % 2 == 1, (n, r)
    assert plan(208) == [16, 13] and len(plan(2016)) == 4 and len(plan(672)) == 3 and plan(1) == []


def test_unlucky_sizes_are_refused():
    for n in (17, 34, 19 * 4, 23 * 3):
        with pytest.raises(RuntimeError, match="not a lucky number"):
            plan(n)
