"""Generates the committed golden fixtures from /root/reference (run in the builder container only;
nothing on the GPU box reads /root/reference).

1. bpti_bv_features.npz -- the real BPTI BV features (53 residues x 500 frames, integer contact
   counts) that survive in the reference tree in HDXer text format
   (notebooks_DEPRECATED/CrossValidation/BPTI/HDXer/BPTI_TFES_RW_bench_R3_k_sequence/train_BPTI_TFES_1/
   {Contacts,Hbonds}_chain_0_res_*.tmp, one file per residue, one line per frame) plus the peptide
   segments of BPTI_residue_segs_trimmed.txt.  Stored as uint8 (lossless).
2. reference_known_answers.json -- the numeric expectations asserted by the reference's own unit
   tests for this path, transcribed with their file:line.

The reference's JAX implementation itself cannot be imported here (no jax/optax/chex in the image),
so there are no reference-generated output vectors; see oracle/bv_oracle.py header.
"""
import glob
import json
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def bpti():
    d = os.path.join(REF, "notebooks_DEPRECATED/CrossValidation/BPTI/HDXer/BPTI_TFES_RW_bench_R3_k_sequence/train_BPTI_TFES_1")
    files = glob.glob(os.path.join(d, "Contacts_chain_0_res_*.tmp"))
    resids = sorted(int(re.search(r"res_(\d+)\.tmp", f).group(1)) for f in files)
    heavy = np.stack([np.loadtxt(os.path.join(d, f"Contacts_chain_0_res_{r}.tmp")) for r in resids])
    acc = np.stack([np.loadtxt(os.path.join(d, f"Hbonds_chain_0_res_{r}.tmp")) for r in resids])
    assert heavy.shape == acc.shape and heavy.shape[1] == 500, heavy.shape
    assert np.all(heavy == np.rint(heavy)) and np.all(acc == np.rint(acc)) and heavy.max() < 256
    segs = np.loadtxt(os.path.join(REF, "notebooks_DEPRECATED/CrossValidation/BPTI/BPTI_residue_segs_trimmed.txt"), dtype=np.int64)
    np.savez_compressed(
        os.path.join(HERE, "bpti_bv_features.npz"),
        heavy=heavy.astype(np.uint8),
        acceptor=acc.astype(np.uint8),
        resids=np.asarray(resids, np.int32),
        segs=segs.astype(np.int32),
    )
    print("bpti", heavy.shape, "max", heavy.max(), acc.max(), "segs", segs.shape)


def known_answers():
    t = "jaxent/tests/unit"
    ka = {
        "bv_log_pf": [
            {"src": f"{t}/models/test_bv_forward_pass.py:94-102", "heavy": [1.0, 2.0, 3.0], "acceptor": [0.5, 1.0, 1.5],
             "bc": 0.35, "bh": 2.0, "expected": [1.35, 2.7, 4.05], "rtol": 1e-6},
            {"src": f"{t}/models/test_bv_forward_pass.py:81-85", "heavy": [0.0, 0.0, 0.0], "acceptor": [0.0, 0.0, 0.0],
             "bc": 0.35, "bh": 2.0, "expected": [0.0, 0.0, 0.0], "rtol": 1e-6},
            {"src": f"{t}/models/test_bv_forward_pass.py:87-92", "heavy": [1.0, 1.0, 1.0], "acceptor": [1.0, 1.0, 1.0],
             "bc": 0.35, "bh": 2.0, "expected": [2.35, 2.35, 2.35], "rtol": 1e-6},
            {"src": f"{t}/models/test_bv_forward_pass.py:119-132", "heavy": [-1.0, 0.0, 1.0], "acceptor": [-0.5, 0.0, 0.5],
             "bc": 0.35, "bh": 2.0, "expected": [-1.35, 0.0, 1.35], "rtol": 1e-6},
        ],
        "bv_uptake": [
            {"src": f"{t}/models/test_bv_forward_pass.py:143-161", "heavy": [0.1], "acceptor": [0.1], "k_ints": [0.1],
             "bc": 0.0, "bh": 0.0, "timepoints": [0.167], "expected_00": float(1.0 - np.exp(-0.1 * 0.167)), "rtol": 1e-4},
            {"src": f"{t}/models/test_bv_forward_pass.py:163-177", "heavy": [100.0], "acceptor": [100.0], "k_ints": [1.0],
             "bc": 0.35, "bh": 2.0, "timepoints": [1.0], "upper_bound_00": 0.1},
            {"src": f"{t}/models/test_bv_forward_pass.py:179-196", "heavy": [[1.0, 1.1], [2.0, 2.1], [3.0, 3.1]],
             "acceptor": [[0.5, 0.6], [1.0, 1.1], [1.5, 1.6]], "k_ints": [0.1, 0.5, 1.0], "bc": 0.35, "bh": 2.0,
             "timepoints": [0.167, 1.0, 10.0], "expected_shape": [3, 3, 2]},
            {"src": f"{t}/models/test_bv_forward_pass.py:198-217", "heavy": [[1.0], [2.0]], "acceptor": [[0.5], [1.0]],
             "k_ints": [0.1, 0.5], "bc": 0.35, "bh": 2.0, "timepoints": [0.001, 0.1, 1.0, 10.0], "monotone_in_time": True},
        ],
        "losses": [
            {"src": f"{t}/opt/test_loss_math.py:47-58", "kind": "l2", "pred": [1.0, 2.0, 3.0], "target": [1.0, 1.0, 1.0],
             "expected": 5.0 / 3.0, "rtol": 1e-5},
            {"src": f"{t}/opt/test_loss_math.py:213-221", "kind": "pf_l2_identity_map", "log_pf": [1.0, 2.0, 3.0],
             "target": [1.0, 2.0, 3.0], "expected": 0.0, "atol": 1e-5},
        ],
        "parity_tolerances": {"src": "jaxent/tests/modules/optimise/test_module_batch_optimise.py:184-193",
                              "loss_rtol": 1e-4, "frame_weights_atol": 1e-4},
    }
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)
    print("known answers written")


if __name__ == "__main__":
    bpti()
    known_answers()
