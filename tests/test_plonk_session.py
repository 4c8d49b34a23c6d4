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
CASES = [(6, 1, 1), (7, 3, 1), (6, 1, 2), (7, 4, 2), (6, 1, 3), (7, 2, 3)]
# (k, seed, circuit variant): 1 = instance + one lookup + two permutation sets,
#                             2 = "wide": two lookups (one two-column), three permutation sets, rotation -1
#                             3 = "phased": two advice phases with a challenge each (running random linear combination, a lookup
#                                 whose input and table are combined with the challenge, a copy from a phase-0 into a phase-1 cell)
SHAPE = {1: "proof_bytes 1216 commitments 14 evals 24", 2: "proof_bytes 1760 commitments 18 evals 37",
         3: "proof_bytes 1184 commitments 14 evals 23"}
GPU_CASES = CASES + [(9, 5, 1), (8, 2, 2)]


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


def parse_poseidon(out):
    """the Poseidon-transcript proof, the exported protocol JSON, the instance values and [tau]G2 the driver printed"""
    lines = out.splitlines()
    proofs = {l.split()[1]: bytes.fromhex(l.split()[2]) for l in lines if l.startswith("poseidon_proof")}
    proto_text = [l for l in lines if l.startswith("protocol_json")][0][len("protocol_json "):]
    instances = [[int.from_bytes(bytes.fromhex(h), "little") for h in l.split()[2:]] for l in lines if l.startswith("instances")]
    w = [int.from_bytes(bytes.fromhex(h), "little") for h in [l for l in lines if l.startswith("s_g2_le")][0].split()[1:]]
    return proofs, proto_text, instances, ((w[0], w[1]), (w[2], w[3]))


@pytest.mark.parametrize("k,seed,variant", [(6, 1, 1), (7, 4, 2), (6, 1, 3)])
def test_poseidon_transcript_proofs_are_accepted_by_the_snark_verifier_model(k, seed, variant, tmp_path):
    """create_proof with snark-verifier's Poseidon transcript + export_protocol_json: the proof is accepted by
    tests/snark_verifier_model.py -- the model that accepts the REFERENCE's shipped chunk and batch proofs
    (tests/test_reference_proofs_kat.py) -- driven by nothing but the exported protocol; tampering is rejected; the exported
    protocol is well-formed for the product's own reader (protocol_json.hpp)."""
    from pairing_model import G2_GEN, g2_neg, pairing_check
    from snark_verifier_model import PoseidonSpec, verify_plonk

    _, out = run("oracle", k, seed, variant)
    proofs, proto_text, instances, s_g2 = parse_poseidon(out)
    proto, proof, spec = json.loads(proto_text), proofs["oracle"], PoseidonSpec()
    decide = lambda lhs, rhs: pairing_check([(lhs, G2_GEN), (rhs, g2_neg(s_g2))])
    lhs, rhs, info = verify_plonk(proto, instances, proof, spec)
    assert decide(lhs, rhs)
    # per phase: advice commitments | challenges (theta joins the last phase), then [m] | [beta, gamma], [z, phi, random] | [y]
    assert (proto["num_witness"], proto["num_challenge"]) == {1: ([3, 1, 4], [1, 2, 1]), 2: ([4, 2, 6], [1, 2, 1]),
                                                             3: ([2, 2, 1, 3], [1, 2, 2, 1])}[variant]
    for pos in (7, len(proof) // 2, len(proof) - 9):
        bad = bytearray(proof)
        bad[pos] ^= 1
        try:
            l2, r2, _ = verify_plonk(proto, instances, bytes(bad), spec)
        except ValueError:
            continue
        assert not decide(l2, r2), pos
    if instances:
        wrong = [instances[0][:]] + instances[1:]
        wrong[0][0] ^= 1
        l3, r3, _ = verify_plonk(proto, wrong, proof, spec)
        assert not decide(l3, r3)
    path = tmp_path / "ours.protocol"
    path.write_text(proto_text)
    import test_protocol_json as TP
    rc, summary = TP.run(str(path))
    assert rc == 0 and summary["k"] == k and summary["domain_matches_mirror"] is True and summary["proof_bytes_shplonk"] == len(proof)


def key(k, seed, variant):
    return f"k{k}_seed{seed}" + ("" if variant == 1 else f"_v{variant}")


@pytest.mark.parametrize("k,seed,variant", CASES)
def test_session_over_the_oracle_verifies_and_matches_the_committed_digest(k, seed, variant):
    proofs, out = run("oracle", k, seed, variant)
    assert "mock_prove: honest witness clean; sabotaged witnesses flagged" in out  # dev::MockProver's check, no proving
    assert SHAPE[variant] in out  # variant 1: 3 advice + m + 2 z + phi + random + 4 h + 2 SHPLONK points
    want = json.load(open(DIGESTS))[key(k, seed, variant)]
    assert hashlib.sha256(proofs["oracle"]).hexdigest() == want


@pytest.mark.gpu
@pytest.mark.parametrize("k,seed,variant", GPU_CASES)
def test_session_on_the_device_gives_identical_proof_bytes(k, seed, variant):
    proofs, out = run("both", k, seed, variant)
    assert "device proof identical to the oracle's" in out
    assert proofs["device"] == proofs["oracle"]
    pos_proofs, _, _, _ = parse_poseidon(out)
    assert pos_proofs["device"] == pos_proofs["oracle"]  # and with snark-verifier's Poseidon transcript as well
    digests = json.load(open(DIGESTS))
    if key(k, seed, variant) in digests:
        assert hashlib.sha256(proofs["device"]).hexdigest() == digests[key(k, seed, variant)]


@pytest.mark.gpu
@pytest.mark.parametrize("k,variant", [(14, 1), (16, 2), (14, 3)])
def test_device_only_session_at_larger_sizes_is_accepted_by_both_verifiers(k, variant):
    """2^14 / 2^16 rows (extended domain 2^16 / 2^18): SRS, keygen and create_proof entirely through the C ABI, Poseidon transcript;
    the pairing-based verifiers (halo2-style and the snark-verifier mirror under the exported protocol) accept, a flipped byte is
    rejected -- no oracle in the loop at this size."""
    r = subprocess.run([binary(), "device", str(k), "3", str(variant)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK") and "accepted by both verifiers" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]
