#!/usr/bin/env python
"""Write the measured parity numbers of the GPU path to a JSON artefact (run on the GPU box):

    python tools/parity_report.py gpurun_out/parity_report.json      -> copied to profiles/rNN_parity_report.json

Per golden case and per output key: max-abs error against the reference's outputs on identical samples and free-running,
the number of rays on the last-sample ReLU kink (excluded), the number of rays flagged as ill-conditioned by the
inverse-CDF (free-running fine pass only), the largest fine-depth difference, and the event / blur tensors formed from OUR
renders against the reference's (train.py:163-177, 299-318) with the dark-pixel population stated explicitly.
Test infrastructure: imports tests/ and oracle/, never shipped.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.cases import CASES                                         # noqa: E402
from tests.test_gpu_render import _run_case, event_tensor_report      # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.json")
    rep = {"tolerance": 1e-4, "device": torch.cuda.get_device_name(0), "cases": {}}
    for name, case in CASES.items():
        entry = {}
        for mode in ("tc", "simt"):
            _, gold, _, rets, r = _run_case(name, mode, inject_z_fine=True)
            entry[f"identical_samples_{mode}"] = r
            if case.n_importance > 0:
                if mode == "tc":
                    entry["image_formation_tc"] = event_tensor_report(case, gold, rets)
                _, _, _, _, r = _run_case(name, mode, inject_z_fine=False)
                entry[f"free_running_{mode}"] = r
        rep["cases"][name] = entry
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(rep, f, indent=1, sort_keys=True)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk.startswith("image")} for k, v in rep["cases"].items()}, indent=1))


if __name__ == "__main__":
    main()
