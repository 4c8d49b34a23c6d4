"""Checks the product's device field header (csrc/ff.cuh) under host emulation against Python big ints.

The PTX carry-chain leaves have a 64-bit-arithmetic emulation when compiled without __CUDA_ARCH__, so
the even/odd accumulator choreography of the Montgomery multiplier is verified here without a GPU;
the -m gpu tests then check the real PTX path against the oracle.
"""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "ff_host.cpp")
SO = os.path.join(HERE, "host_emul", "libff_host.so")
R = 1 << 256
MODS = [0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
        0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47]


@pytest.fixture(scope="module")
def lib():
    hdr = os.path.join(HERE, "..", "scroll-prover_b200", "csrc", "ff.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    return ctypes.CDLL(SO)


def tolimbs(vs):
    return np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for v in vs], dtype=np.uint32)


def toints(a):
    return [sum(int(x) << (32 * i) for i, x in enumerate(row)) for row in a]


@pytest.mark.parametrize("field", [0, 1])
def test_ff_host_emulation_matches_bigint(lib, field):
    p = MODS[field]
    rng = random.Random(1 + field)
    edge = [0, 1, 2, p - 1, p - 2, R % p, (p - 1) // 2, 1 << 253, p - 3, 0xFFFFFFFF, (1 << 224) - 1,
            (1 << 32), (1 << 64) - 1, p - (1 << 32)]
    vals = edge + [rng.randrange(p) for _ in range(2000)]
    n = 20000
    A = [rng.choice(vals) for _ in range(n)]
    B = [rng.choice(vals) for _ in range(n)]
    A[: len(edge) ** 2] = [x for x in edge for _ in edge]
    B[: len(edge) ** 2] = [y for _ in edge for y in edge]
    a, b = tolimbs(A), tolimbs(B)
    r = np.zeros_like(a)
    vp = ctypes.c_void_p

    def run(op, cnt=n):
        lib.ff_host_op(field, op, r.ctypes.data_as(vp), a.ctypes.data_as(vp), b.ctypes.data_as(vp), ctypes.c_uint64(cnt))
        return toints(r[:cnt])

    rinv = pow(R, -1, p)
    assert run(0) == [x * y * rinv % p for x, y in zip(A, B)]
    assert run(1) == [(x + y) % p for x, y in zip(A, B)]
    assert run(2) == [(x - y) % p for x, y in zip(A, B)]
    assert run(4) == [x * rinv % p for x in A]
    assert run(5) == [x * R % p for x in A]
    assert run(6) == [(-x) % p for x in A]
    assert run(7) == [x * x * rinv % p for x in A]
    assert run(3, 60) == [(pow(x * rinv % p, -1, p) * R % p if x else 0) for x in A[:60]]


def test_ec_host_emulation_matches_oracle(lib):
    """csrc/ec.cuh XYZZ formulas (incl. doubling / inverse / identity branches) vs the C oracle."""
    from oracle import oracle as O

    vp = ctypes.c_void_p
    pts = O.fill_points(6, 99, 2)
    G = O.g1_generator()
    ident = np.zeros(8, np.uint64)
    negG = np.concatenate([G[:4], O.fq_sub(np.zeros(4, np.uint64), G[4:])])
    affs = [ident, G, negG] + [p for p in pts]
    jacs = []
    for a in affs:
        j = O.g1_from_affine(a)
        jacs.append(j)
        jacs.append(O.g1_double(O.g1_add(j, O.g1_from_affine(G))))  # non-trivial Z
    for j in jacs:
        for a in affs:
            for op in (0, 1, 2):
                out = np.zeros(8, np.uint64)
                jj = np.ascontiguousarray(j)
                aa = np.ascontiguousarray(a)
                lib.ec_host_op(op, out.ctypes.data_as(vp), jj.ctypes.data_as(vp), aa.ctypes.data_as(vp))
                exp = O.g1_double(jj) if op == 2 else O.g1_add_mixed(jj, aa)
                assert np.array_equal(out, O.g1_to_affine(exp)), (op,)
    # acc == q (doubling branch) and acc == -q (identity branch) with non-trivial Z
    for a in pts[:3]:
        j = O.g1_add(O.g1_double(O.g1_from_affine(a)), O.g1_from_affine(np.concatenate([a[:4], O.fq_sub(np.zeros(4, np.uint64), a[4:])])))
        for op in (0, 1):
            for q in (a, np.concatenate([a[:4], O.fq_sub(np.zeros(4, np.uint64), a[4:])])):
                out = np.zeros(8, np.uint64)
                qq = np.ascontiguousarray(q)
                lib.ec_host_op(op, out.ctypes.data_as(vp), j.ctypes.data_as(vp), qq.ctypes.data_as(vp))
                assert np.array_equal(out, O.g1_to_affine(O.g1_add_mixed(j, qq)))
