"""Execute the reference's two entry points (train_gan.py per --phase, train_segmenter.py) unmodified, with the model modules
(`adversarial`, `source_segmenter`) replaced by recorders, and commit what they would have passed to Full_DRN / Trainer / train
as tests/golden/reference_config.json.  This pins the configuration dictionaries and the per-phase overrides of
train_gan.py:24-129 and train_segmenter.py:22-77 to the reference's own code (including the NameError that makes
`--phase fine-tune` unusable, train_gan.py:121).

    python tests/golden/make_reference_config_vectors.py        # needs /root/reference; the tests only read the .json
"""
import copy
import importlib.util
import json
import os
import sys
import tempfile
import types

REF = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_config.json")


def _recorder_module(name, log):
    m = types.ModuleType(name)

    class Full_DRN(object):
        def __init__(self, **kw):
            log["Full_DRN"] = copy.deepcopy(kw)

    class Trainer(object):
        def __init__(self, net, *lists, **kw):
            kw = dict(kw)
            rec = {}
            for k, v in list(kw.items()):
                if k.endswith("_list"):
                    rec[k + "_len"] = None if v is None else len(v)
                else:
                    rec[k] = copy.deepcopy(v)
            rec["positional_lists_len"] = [None if v is None else len(v) for v in lists]
            log["Trainer"] = rec

        def train(self, **kw):
            log["train"] = copy.deepcopy(kw)
    m.Full_DRN, m.Trainer = Full_DRN, Trainer
    return m


def _run(script, model_module, call):
    """fresh execution of one entry-point file; returns what it handed to the (recorded) model module"""
    log = {}
    tf = types.ModuleType("tensorflow")
    tf.python = types.ModuleType("tensorflow.python")
    tf.python.debug = types.ModuleType("tensorflow.python.debug")
    stubs = {"tensorflow": tf, "tensorflow.python": tf.python, "tensorflow.python.debug": tf.python.debug,
             "nibabel": types.ModuleType("nibabel"), model_module: _recorder_module(model_module, log)}
    saved = {k: sys.modules.get(k) for k in list(stubs) + ["lib"]}
    sys.modules.update(stubs)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="pnp_cfg_")
    os.symlink(os.path.join(REF, "lists"), os.path.join(tmp, "lists"))      # the scripts read ./lists/* relative to the cwd
    os.chdir(tmp)                                                            # ... and write ./general_log, ./tmp_exps there
    real_system = os.system
    os.system = lambda cmd: log.setdefault("os_system", []).append(cmd) or 0   # train_segmenter.py:69-70 launches tensorboard
    env_before = os.environ.get("CUDA_VISIBLE_DEVICES")
    try:
        sys.path.insert(0, REF)                                               # `from lib import _read_lists`
        spec = importlib.util.spec_from_file_location("pnp_reference_" + script, os.path.join(REF, script + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        try:
            call(mod)
        except Exception as e:                                                # noqa: BLE001 -- the reference's own failure is the datum
            log["error"] = "%s: %s" % (type(e).__name__, e)
        log["CUDA_VISIBLE_DEVICES_set_by_script"] = os.environ.get("CUDA_VISIBLE_DEVICES")
    finally:
        os.system = real_system
        os.chdir(cwd)
        sys.path.remove(REF)
        sys.modules.pop("lib", None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if env_before is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = env_before
    return log


def main():
    out = {"train_gan": {}, "train_segmenter": None}
    for phase in ("pre-train", "train-gan", "fine-tune", "bogus"):
        out["train_gan"][phase] = _run("train_gan", "adversarial", lambda m, p=phase: m.main(phase=p))
    out["train_segmenter"] = _run("train_segmenter", "source_segmenter", lambda m: m.main())
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT)
    for ph, v in out["train_gan"].items():
        print(ph, sorted(v.keys()), v.get("error"))
    print("train_segmenter", sorted(out["train_segmenter"].keys()))


if __name__ == "__main__":
    main()
