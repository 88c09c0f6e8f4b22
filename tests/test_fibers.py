"""CPU suite: the batch producer's worker scheduler and context switch (lamsa_b200/csrc/fiber_sched.cu,
fiber_switch.cpp) without a GPU: workers that keep floating-point and stack state across thousands of
switches, on one and on several scheduler threads."""
import ctypes as C

import pytest

import lamsa_b200


@pytest.mark.parametrize("n,yields,threads", [(1, 10, 1), (64, 200, 1), (1000, 50, 4), (37, 1000, 3), (5, 0, 2)])
def test_fiber_selftest(n, yields, threads):
    lib = lamsa_b200.load_library()
    lib.lb2_fiber_selftest.argtypes = [C.c_int, C.c_int, C.c_int]
    assert lib.lb2_fiber_selftest(n, yields, threads) == 0
