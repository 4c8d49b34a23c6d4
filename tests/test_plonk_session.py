"""The create_proof / verify_proof session of scroll-prover_b200/plonk_b200.hpp (reference entry
/root/reference/integration/src/prove.rs:37-39 gen_halo2_chunk_proof, check :50-53 verify_chunk_proof).

CPU (no device): the prover runs over the CPU oracle (tests/cpp/oracle_ops.hpp), the proof must verify under the host pairing
verifier, every tampering must be rejected (asserted inside tests/cpp/test_plonk_session.cpp), and the proof bytes must hash to
the COMMITTED digest (tests/golden/plonk_session_digests.json) -- so neither the oracle nor the prover can drift unnoticed.
GPU: the same prover over the CUDA path through the C ABI must give IDENTICAL PROOF BYTES (and therefore the same digest).
"""
import hashlib
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_plonk_session.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "test_plonk_session")
DIGESTS = os.path.join(ROOT, "tests", "golden", "plonk_session_digests.json")
CASES = [(6, 1, 1), (7, 3, 1), (6, 1, 2), (7, 4, 2)]  # (k, seed, circuit variant: 1 = instance + one lookup + two permutation sets,
#                                                          2 = "wide": two lookups (one two-column), three permutation sets, rotation -1)
SHAPE = {1: "proof_bytes 1216 commitments 14 evals 24", 2: "proof_bytes 1760 commitments 18 evals 37"}


def binary():
    deps = [SRC, os.path.join(ROOT, "tests", "cpp", "oracle_ops.hpp")] + [os.path.join(ROOT, "scroll-prover_b200", h) for h in
                                                                         ("plonk_b200.hpp", "halo2_b200.hpp", "pairing_bn254.hpp", "serde_bn254.hpp")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        lib, orc = os.path.join(ROOT, "scroll-prover_b200"), os.path.join(ROOT, "oracle")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", BIN, SRC, "-L" + lib, "-lb200zk", "-Wl,-rpath," + lib, "-L" + orc, "-loracle",
                               "-Wl,-rpath," + orc])
    return BIN


def run(mode, k, seed, variant=1):
    r = subprocess.run([binary(), mode, str(k), str(seed), str(variant)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-2000:] + r.stderr[-2000:]
    proofs = {l.split()[1]: bytes.fromhex(l.split()[2]) for l in r.stdout.splitlines() if l.startswith("proof_sha_input")}
    return proofs, r.stdout


def test_blake2b_transcript_hash_matches_hashlib():
    for msg in (b"", b"\x01" + bytes(range(64)), bytes(range(256)) * 3 + b"tail"):
        out = subprocess.run([binary(), "blake2b", msg.hex()], capture_output=True, text=True, timeout=60).stdout.strip()
        assert out == hashlib.blake2b(msg, digest_size=64, person=b"Halo2-Transcript").hexdigest()


def key(k, seed, variant):
    return f"k{k}_seed{seed}" + ("" if variant == 1 else f"_v{variant}")


@pytest.mark.parametrize("k,seed,variant", CASES)
def test_session_over_the_oracle_verifies_and_matches_the_committed_digest(k, seed, variant):
    proofs, out = run("oracle", k, seed, variant)
    assert SHAPE[variant] in out  # variant 1: 3 advice + m + 2 z + phi + random + 4 h + 2 SHPLONK points
    want = json.load(open(DIGESTS))[key(k, seed, variant)]
    assert hashlib.sha256(proofs["oracle"]).hexdigest() == want


@pytest.mark.gpu
@pytest.mark.parametrize("k,seed,variant", CASES + [(9, 5, 1), (8, 2, 2)])
def test_session_on_the_device_gives_identical_proof_bytes(k, seed, variant):
    proofs, out = run("both", k, seed, variant)
    assert "device proof identical to the oracle's" in out
    assert proofs["device"] == proofs["oracle"]
    digests = json.load(open(DIGESTS))
    if key(k, seed, variant) in digests:
        assert hashlib.sha256(proofs["device"]).hexdigest() == digests[key(k, seed, variant)]
