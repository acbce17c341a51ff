"""Generate tests/golden/reference_eval_vectors.json by EXECUTING the reference's own test protocol in this container.

`adversarial.py` cannot be imported (TensorFlow 1.4 is absent), but `Trainer.test_eval` (adversarial.py:993-1052) and
`Trainer.sample_metric_stddev` (:1054-1084) are plain numpy around one `sess.run`.  This script pulls exactly those two function
definitions out of /root/reference/adversarial.py with `ast` (nothing is copied into the repository), compiles them unmodified and
runs them with

  * `read_nii_image`, `_label_decomp`, `_dice`, `_jaccard` from the reference's own lib.py (imported with the TF / nibabel shims of
    make_reference_lib_vectors.py); `read_nii_image` itself is replaced by a lookup into seeded in-memory subjects, because nibabel is
    absent -- the NIfTI reader is pinned separately, by byte-level known answers (tests/test_nifti_eval_cpu.py),
  * the module constants the functions read (`raw_size`, `label_size`, `contour_map`, `floor`) taken from the reference source,
  * a stand-in session whose `run` answers the two fetches (`compact_pred`, `confusion_matrix`) with a fixed, network-free
    "segmenter": pred = clip(floor(1.5 * middle channel + 2), 0, 4), and the confusion matrix of argmax(one-hot labels) vs pred.

What is recorded is therefore the reference's PROTOCOL -- flip, frame list, shuffling through the global numpy RNG, how many batches,
which rows of a short last batch stay zero and are still counted, the per-subject Dice / Jaccard, the summed confusion matrix and the
(Dice list, `subject_level_list[:1]`) pair that `sample_metric_stddev` returns.  tests/test_nifti_eval_cpu.py runs the product's
`evaluation.run_test_eval` with the same stand-in predictor on the same seeded subjects and must reproduce every number.

    python tests/golden/make_reference_eval_vectors.py          # needs /root/reference; the tests only read the .json
"""
import ast
import importlib.util
import json
import logging
import math
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "reference_eval_vectors.json")

CASES = [  # (name, batch_size, depths of the subjects, flip_correction, numpy seed set right before test_eval)
    ("b2_d7_d6_flip", 2, [7, 6], True, 11),
    ("b4_d9_noflip", 4, [9], False, 5),
    ("b3_d8_d5_d12_flip", 3, [8, 5, 12], True, 2024),
]


def make_subject(seed, depth):
    """[256, 256, depth] float32 image and an integer label volume correlated with it (shared with the test)"""
    rng = np.random.RandomState(seed)
    raw = rng.standard_normal((256, 256, depth)).astype(np.float32)
    noise = rng.standard_normal((256, 256, depth)) * 0.4
    raw_y = np.clip(np.floor(1.5 * raw.astype(np.float64) + 2.0 + noise), 0, 4).astype(np.int16)
    return raw, raw_y


def stand_in_prediction(vol):
    """the network-free "segmenter" (shared with the test): labels from the middle channel of each row of the batch"""
    v = np.asarray(vol, np.float64)[..., 1]
    return np.clip(np.floor(1.5 * v + 2.0), 0, 4).astype(np.int64)


def _reference_functions():
    sys.path.insert(0, HERE)
    import make_reference_lib_vectors as shim
    sys.modules["tensorflow"] = shim._make_tf_shim()
    sys.modules["nibabel"] = types.ModuleType("nibabel")
    spec = importlib.util.spec_from_file_location("pnp_reference_lib", os.path.join(REF, "lib.py"))
    lib = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lib)
    src = open(os.path.join(REF, "adversarial.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "os": os, "logging": logging, "floor": math.floor, "_label_decomp": lib._label_decomp, "_dice": lib._dice,
          "_jaccard": lib._jaccard}
    wanted_consts = {"contour_map", "raw_size", "label_size"}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and getattr(node.targets[0], "id", None) in wanted_consts:
            ns[node.targets[0].id] = ast.literal_eval(node.value)
        if isinstance(node, ast.ClassDef) and node.name == "Trainer":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in ("test_eval", "sample_metric_stddev"):
                    mod = ast.Module(body=[fn], type_ignores=[])
                    exec(compile(mod, os.path.join(REF, "adversarial.py"), "exec"), ns)
    assert wanted_consts <= set(ns) and "test_eval" in ns and "sample_metric_stddev" in ns
    return ns


class _Session(object):
    def __init__(self, net, num_cls):
        self.net, self.num_cls, self.calls = net, num_cls, []

    def run(self, fetches, feed_dict):
        assert fetches == [self.net.compact_pred, self.net.confusion_matrix]
        assert feed_dict[self.net.keep_prob] == 1.0 and feed_dict[self.net.mr_front_bn] is False and feed_dict[self.net.ct_front_bn] is False
        vol, vol_y = feed_dict[self.net.ct], feed_dict[self.net.ct_y]
        pred = stand_in_prediction(vol)
        lab = np.argmax(vol_y, axis=-1)
        cm = np.zeros((self.num_cls, self.num_cls), np.int64)
        np.add.at(cm, (lab.reshape(-1), pred.reshape(-1)), 1)          # tf.confusion_matrix: rows = labels, columns = predictions
        self.calls.append(int(np.count_nonzero(np.abs(vol).reshape(vol.shape[0], -1).sum(1))))
        return pred, cm


def main():
    ns = _reference_functions()
    out = {"generator": "tests/golden/make_reference_eval_vectors.py", "source": "adversarial.py:993-1084 executed", "cases": {}}
    for name, B, depths, flip, seed in CASES:
        volumes = {}
        tmp = tempfile.mkdtemp()
        nii, lab = [], []
        for i, d in enumerate(depths):
            raw, raw_y = make_subject(1000 * seed + i, d)
            fi, fl = os.path.join(tmp, "img_%d.nii" % i), os.path.join(tmp, "lab_%d.nii" % i)
            open(fi, "w").close()                        # test_eval checks os.path.isfile(nii_fid)
            volumes[fi], volumes[fl] = raw, raw_y
            nii.append(fi)
            lab.append(fl)
        ns["read_nii_image"] = lambda fid: volumes[fid]
        net = types.SimpleNamespace(batch_size=B, compact_pred="compact_pred", confusion_matrix="confusion_matrix", ct="ct", ct_y="ct_y",
                                    keep_prob="keep_prob", mr_front_bn="mr_front_bn", ct_front_bn="ct_front_bn")
        me = types.SimpleNamespace(num_cls=5, net=net, test_label_list=lab, test_nii_list=nii)
        me.sample_metric_stddev = lambda lst, me=me: ns["sample_metric_stddev"](me, lst)
        sess = _Session(net, 5)
        outdir = tempfile.mkdtemp()
        np.random.seed(seed)
        dice_list, jac_quirk = ns["test_eval"](me, sess, outdir, flip_correction=flip)
        all_cm = np.loadtxt(os.path.join(outdir, "cm.csv"))
        out["cases"][name] = {
            "batch_size": B, "depths": depths, "flip_correction": flip, "seed": seed,
            "subject_dice_list": [float(v) for v in dice_list],
            "subject_jaccard_quirk": np.asarray(jac_quirk, np.float64).tolist(),        # shape [1, 2]: the reference's `[:1]` slip
            "all_cm": all_cm.tolist(),
            "forward_calls": len(sess.calls), "nonzero_rows_per_call": sess.calls,
        }
        print(name, "calls", len(sess.calls), "dice", np.round(dice_list, 4))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
