"""Extracts the small parity-pinning fixtures from the read-only reference checkout into
tests/golden/reference_fixtures.json.  Run in the build container:
    python tests/golden/make_golden.py
/root/reference does not exist on the GPU box, so tests only ever read the JSON written here.
"""
import base64
import hashlib
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")


def main():
    out = {"source": "scroll-tech/scroll-prover @ ecba83e", "files": {}}
    proto = json.load(open(f"{REF}/release-v0.13.1/chunk.protocol"))
    out["chunk_protocol"] = {
        "path": "release-v0.13.1/chunk.protocol",
        "domain": proto["domain"],
        "preprocessed": proto["preprocessed"],
        "num_witness": proto["num_witness"],
        "quotient_num_chunk": proto["quotient"]["num_chunk"],
    }
    for name in ("vk_chunk.vkey", "vk_batch.vkey", "vk_bundle.vkey"):
        b = open(f"{REF}/release-v0.13.1/{name}", "rb").read()
        out["files"][name] = {"path": f"release-v0.13.1/{name}", "sha256": hashlib.sha256(b).hexdigest(), "hex": b.hex()}
    yul = open(f"{REF}/release-v0.13.1/evm_verifier.yul").read().splitlines()
    out["evm_verifier_yul"] = {
        "path": "release-v0.13.1/evm_verifier.yul",
        "line17": yul[16].strip(),
        "line18": yul[17].strip(),
        "lines1230_1240": [l.strip() for l in yul[1229:1240]],
    }
    pd = open(f"{REF}/release-v0.13.1/proof.data", "rb").read()
    out["files"]["proof.data"] = {"path": "release-v0.13.1/proof.data", "sha256": hashlib.sha256(pd).hexdigest(),
                                  "len": len(pd), "accumulator_hex": pd[:384].hex()}
    fp = json.load(open(f"{REF}/integration/tests/test_data/full_proof_1.json"))["chunk_proofs"][0]
    proof = fp["proof"]
    pb = base64.b64decode(proof) if isinstance(proof, str) else bytes(proof)
    inst = fp["instances"]
    ib = base64.b64decode(inst) if isinstance(inst, str) else bytes(inst)
    out["full_proof_1"] = {"path": "integration/tests/test_data/full_proof_1.json", "proof_len": len(pb),
                           "proof_hex": pb.hex(), "instances_len": len(ib), "instances_accumulator_hex": ib[:384].hex(),
                           "vk_hex": base64.b64decode(fp["vk"]).hex()}
    cfg = {}
    for i in range(1, 7):
        cfg[f"layer{i}"] = json.load(open(f"{REF}/integration/configs/layer{i}.config"))
    out["layer_configs"] = cfg
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
