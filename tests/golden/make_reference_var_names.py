"""Generates tests/golden/reference_var_names.json from the reference's own variable-name lists
(/root/reference/lists/*): the checkpoint naming contract that both the oracle and the product must honour.
Run in the build container (the reference tree is not available on the GPU box)."""
import json
import os

REF = "/root/reference/lists"
out = {}
for f in ("half_zip_mri_vars", "half_zip_ct_vars", "old_bn_list", "pred_bn_list"):
    with open(os.path.join(REF, f)) as fd:
        out[f] = [l.strip()[:-2] if l.strip().endswith(":0") else l.strip() for l in fd if len(l.strip()) >= 3]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_var_names.json"), "w") as fd:
    json.dump(out, fd, indent=0)
print({k: len(v) for k, v in out.items()})
