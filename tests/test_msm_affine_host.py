"""The experimental batched-affine bucket accumulation (scroll-prover_b200/csrc/msm_affine.cuh; off by default on the GPU,
B200ZK_MSM_AFFINE=1) under host emulation: its per-thread kernel bodies, run thread by thread and orchestrated like
msm_affine_accumulate, must produce exactly the bucket sums of the XYZZ mixed-addition path -- including the cases the
affine formulas do not cover by themselves: equal points (tangent), opposite points (cancellation to the identity, and
continuing afterwards), identity bases, empty / single / odd-sized / very large buckets, chunk boundaries inside buckets.
CPU only."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "msm_affine_host.cpp")
SO = os.path.join(HERE, "host_emul", "libmsm_affine_host.so")
HDRS = [os.path.join(HERE, "..", "scroll-prover_b200", "csrc", h) for h in ("msm_affine.cuh", "ec.cuh", "ff.cuh")]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(p) for p in [SRC] + HDRS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    return C.CDLL(SO)


def run(lib, bases, buckets, L):
    """buckets: list of lists of (base index, negate) -> (affine sums by the tree, by the reference), each (NB, 8) u64"""
    entries = np.array([i | (0x80000000 if neg else 0) for b in buckets for i, neg in b] or [0], dtype=np.uint32)
    offsets = np.zeros(len(buckets) + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum([len(b) for b in buckets])
    nb = len(buckets)
    m = int(offsets[-1])
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    got = np.zeros((nb, 8), np.uint64)
    want = np.zeros((nb, 8), np.uint64)
    vp = C.c_void_p
    levels = lib.msm_affine_host(vp(bases.ctypes.data), vp(entries.ctypes.data), vp(offsets.ctypes.data), C.c_uint64(nb),
                                 C.c_uint64(max(m, 1)), C.c_uint32(L), vp(got.ctypes.data))
    lib.msm_buckets_reference(vp(bases.ctypes.data), vp(entries.ctypes.data), vp(offsets.ctypes.data), C.c_uint64(nb), vp(want.ctypes.data))
    return got, want, levels


@pytest.mark.parametrize("L", [1, 3, 16])
def test_random_buckets_of_every_size(lib, L):
    rng = random.Random(100 + L)
    bases = O.fill_points(200, 0xAFF1, 4)
    sizes = [0, 1, 2, 3, 0, 5, 8, 17, 64, 1, 0, 0, 33, 2, 129, 7, 1000, 0, 4, 31]
    buckets = [[(rng.randrange(200), rng.random() < 0.5) for _ in range(s)] for s in sizes]
    got, want, levels = run(lib, bases, buckets, L)
    assert np.array_equal(got, want)
    assert levels == 11  # ceil(log2(total entries = 1307))


def test_equal_points_take_the_tangent_and_opposite_points_cancel(lib):
    bases = O.fill_points(8, 0xD0B1, 2)
    z = np.zeros(8, np.uint64)
    bases = np.vstack([bases, z[None, :]])  # index 8: the identity as a base
    P, Q = 0, 1
    buckets = [
        [(P, False), (P, False)],                                  # P + P: tangent
        [(P, False), (P, True)],                                   # P - P: identity
        [(P, False), (P, True), (Q, False)],                       # cancellation, then a leftover
        [(P, False), (P, False), (P, False), (P, False), (P, False)],  # 5P through two doublings and an addition
        [(P, True), (Q, False), (P, False), (Q, True)],            # (−P + Q) + (P − Q) = identity at the second level
        [(8, False), (P, False)],                                  # identity base first
        [(P, False), (8, True)],                                   # identity base second (the sign of (0,0) is ignored)
        [(8, False), (8, False), (8, False)],                      # only identities
        [(P, False)] * 64,                                         # 64 P: six levels of pure doublings
        [(Q, False), (Q, False), (Q, True), (Q, True), (Q, False)],
    ]
    for L in (1, 2, 16):
        got, want, _ = run(lib, bases, buckets, L)
        assert np.array_equal(got, want), L
    assert not got[1].any() and not got[4].any() and not got[7].any()


def test_a_giant_bucket_next_to_many_small_ones(lib):
    """the shape of real witness columns: one digit value repeated thousands of times, the rest sparse; chunk boundaries of the
    per-thread slots fall inside the giant bucket at every level"""
    rng = random.Random(7)
    bases = O.fill_points(64, 0xB16, 4)
    buckets = [[(rng.randrange(64), rng.random() < 0.3) for _ in range(rng.choice([0, 0, 1, 2]))] for _ in range(300)]
    buckets[17] = [(rng.randrange(64), rng.random() < 0.5) for _ in range(5000)]
    buckets[299] = [(rng.randrange(4), False) for _ in range(777)]  # few distinct bases: many doublings
    got, want, levels = run(lib, bases, buckets, 16)
    assert np.array_equal(got, want)
    assert levels == 13


def test_no_entries_at_all(lib):
    bases = O.fill_points(4, 1, 1)
    got, want, _ = run(lib, bases, [[], [], []], 16)
    assert not got.any() and np.array_equal(got, want)


# ---- the same pipeline with the REAL kernel source under the CUDA emulation layer (tests/host_emul/cuda_emu.hpp): the
# offset-scan kernels run as blocks of 256 threads meeting at barriers, the activity flag gates pass A and the inversion
EMU_SRC = os.path.join(HERE, "host_emul", "msm_affine_emu.cpp")
EMU_SO = os.path.join(HERE, "host_emul", "libmsm_affine_emu.so")
EMU_HDRS = HDRS + [os.path.join(HERE, "..", "scroll-prover_b200", "csrc", "msm_affine_kernels.cuh"), os.path.join(HERE, "host_emul", "cuda_emu.hpp")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(p) for p in [EMU_SRC] + EMU_HDRS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", EMU_SO, EMU_SRC])
    return C.CDLL(EMU_SO)


def run_emu(emu, lib, bases, buckets, L):
    entries = np.array([i | (0x80000000 if neg else 0) for b in buckets for i, neg in b] or [0], dtype=np.uint32)
    offsets = np.zeros(len(buckets) + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum([len(b) for b in buckets])
    nb, m = len(buckets), int(offsets[-1])
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    got, want = np.zeros((nb, 8), np.uint64), np.zeros((nb, 8), np.uint64)
    vp = C.c_void_p
    active = emu.msm_affine_emu(vp(bases.ctypes.data), vp(entries.ctypes.data), vp(offsets.ctypes.data), C.c_uint64(nb), C.c_uint64(max(m, 1)),
                                C.c_uint32(L), vp(got.ctypes.data))
    lib.msm_buckets_reference(vp(bases.ctypes.data), vp(entries.ctypes.data), vp(offsets.ctypes.data), C.c_uint64(nb), vp(want.ctypes.data))
    return got, want, active


def test_emulated_kernels_small_mixed_buckets(emu, lib):
    rng = random.Random(31)
    bases = O.fill_points(50, 0xE31, 4)
    sizes = [0, 1, 2, 3, 0, 5, 8, 17, 64, 1, 0, 33, 130, 7]
    buckets = [[(rng.randrange(50), rng.random() < 0.5) for _ in range(s)] for s in sizes]
    buckets[3] = [(0, False), (0, False), (0, True)]  # tangent, then cancellation against the copy
    got, want, active = run_emu(emu, lib, bases, buckets, 16)
    assert np.array_equal(got, want)
    assert active == 8  # the largest bucket (130 entries) needs ceil(log2 130) levels; the later ones are skipped by the flag


def test_emulated_kernels_many_buckets_span_several_scan_tiles(emu, lib):
    """5000 buckets (three 2048-bucket scan tiles), mostly empty or tiny, one large: the tile sums / tile offsets / in-tile
    offsets of the real scan kernels must line every bucket up"""
    rng = random.Random(32)
    bases = O.fill_points(40, 0xE32, 4)
    buckets = [[(rng.randrange(40), rng.random() < 0.4) for _ in range(rng.choice([0, 0, 0, 1, 1, 2, 3]))] for _ in range(5000)]
    buckets[4321] = [(rng.randrange(40), False) for _ in range(300)]
    got, want, active = run_emu(emu, lib, bases, buckets, 16)
    assert np.array_equal(got, want)
    assert active == 9
