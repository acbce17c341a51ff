"""SURVEY 8(f) rows on the device:
  f1  the phase hand-over chain  train_segmenter.py -> checkpoint -> train_gan.py --phase pre-train
      (reference: train_gan.py:74-77, adversarial.py:796-801 = restore(no_gan) -> _load_batch_norm_weights -> _adapt_copy_weights,
      :503-531, :706-765) executed end to end on the GPU and compared with the oracle, whose side of the transplant is done
      independently in numpy from the reference's own name lists (tests/golden/reference_var_names.json = lists/half_zip_*_vars,
      lists/old_bn_list, lists/pred_bn_list);
  f3  the evaluation path (adversarial.py:894-922 / 993-1052: inference-mode BN -- folded into the tcgen05 epilogue here --,
      keep_prob 1, hard Dice + confusion matrix) against the oracle.
"""
import json
import os

import numpy as np
import pytest
import torch

from tests.util import check
from tests.test_parity_configs_gpu import loss_close, state_close, _bn_noise

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))
B = 2


def reference_transplant(ckpt, P, gold):
    """what adversarial.py:503-531 (no_gan), :743-765 and :706-741 do to a freshly initialised GAN graph, on plain dicts"""
    out = dict(P)
    for k, v in ckpt.items():                                                    # restore(no_gan=True)
        if k in out and not any(s in k for s in ("adapt", "cls", "Adam")) and ("group" in k or "output" in k):
            out[k] = np.asarray(v)
    for old, new in zip(gold["old_bn_list"], gold["pred_bn_list"]):              # _load_batch_norm_weights
        new = new.split(":")[0]
        out["group_%s/%s" % (new.split("_")[1], new)] = np.asarray(ckpt[old.split(":")[0]])
    for mr, ad in zip(gold["half_zip_mri_vars"], gold["half_zip_ct_vars"]):      # _adapt_copy_weights
        out[ad.split(":")[0]] = out[mr.split(":")[0]]
    return out


def test_segmenter_checkpoint_hands_over_to_pretrain_discriminator_step(tmp_path):
    import pnp_b200  # noqa: F401
    from pnp_b200 import runtime as rt, adversarial as adv, source_segmenter as seg
    from pnp_b200.train_gan import configure
    from oracle.pnp_graphs import (OracleAdversarial, OracleSegmenter, init_numpy_params, synthetic_images, synthetic_labels)
    rt.set_conv_backend("auto")
    gold = json.load(open(os.path.join(HERE, "golden", "reference_var_names.json")))
    # ---- phase 0: source segmenter, two Adam steps on the GPU, checkpoint --------------------------------------------
    ws, bns = OracleSegmenter.layout()
    Pseg = init_numpy_params(ws, bns, 0, 0.05)
    net = seg.Full_DRN(channels=3, n_class=5, batch_size=B,
                       cost_kwargs={"cross_flag": True, "miu_cross": 1.0, "dice_flag": True, "miu_dice": 1.0, "regularizer": 1e-4})
    rt.load_state_dict(Pseg)
    tr = seg.Trainer(net, [], [], num_cls=5, batch_size=B, optimizer="adam", opt_kwargs={"learning_rate": 1e-3})
    x, lab = synthetic_images(B, 1234), synthetic_labels(B, 99)
    xg, yg = tr.feed(x, torch.from_numpy(lab))
    for _ in range(2):
        tr.train_step(xg, yg, keep_prob=1.0)
    out_dir = str(tmp_path / "seg")
    os.makedirs(out_dir)
    tr.save(os.path.join(out_dir, "model.cpkt"), out_dir)
    ck_path = os.path.join(out_dir, "latest.npz")
    ckpt = dict(np.load(ck_path))
    assert "group_3/Variable_1/Adam" in ckpt and "pnp/global_step" in ckpt and int(ckpt["pnp/global_step"]) == 2
    assert not np.array_equal(ckpt["BatchNorm_5/moving_mean"], np.zeros_like(ckpt["BatchNorm_5/moving_mean"]))
    # ---- phase 1: pre-train graph, the reference's restore chain ------------------------------------------------------
    ws, bns = OracleAdversarial.layout()
    Padv = init_numpy_params(ws, bns, 1, 0.05)            # a DIFFERENT seed: everything the chain must overwrite differs
    _bn_noise(Padv, bns, 11)
    for n, s in ws:
        if "cls" in n:
            Padv[n] = np.clip(Padv[n] * 0.5, -0.05, 0.05).astype(np.float32)
    ck, nc, tc = configure("pre-train")
    anet = adv.Full_DRN(channels=3, n_class=5, batch_size=B, cost_kwargs=ck, network_config=nc, critic_keep_prob=1.0)
    rt.load_state_dict(Padv)
    atr = adv.Trainer(anet, num_cls=5, batch_size=B, opt_kwargs={"learning_rate": 3e-4}, train_config=tc)
    anet.restore(ck_path, no_gan=True)
    anet.load_batch_norm_weights(ck_path)
    anet.adapt_copy_weights()
    want = reference_transplant(ckpt, Padv, gold)
    got = rt.state_dict()
    moved = 0
    for n in want:
        assert np.array_equal(got[n], want[n]), "transplant mismatch at %s" % n
        moved += int(not np.array_equal(want[n], Padv[n]))
    print("  transplant: %d of %d variables overwritten from the segmenter checkpoint, all bit-identical to the reference's mapping" % (moved, len(want)))
    assert moved >= 33 + 120 + 101
    # ---- the first pre-train D step on the transplanted state vs the oracle on the same state --------------------------
    oracle = OracleAdversarial(want, B, lambda_mask_loss=0, dis_sub_iter=tc["dis_sub_iter"], gen_sub_iter=1, critic_keep_prob=1.0)
    mr, ct = synthetic_images(B, 1234), synthetic_images(B, 4321, 0.3, 0.8)
    ro = oracle.d_step(mr, ct, keep_prob=1.0)
    terms = atr.d_step(mr.to(DEV), ct.to(DEV), keep_prob=1.0)
    loss_close("dis_loss after hand-over", atr.loss_value(terms), ro["dis_loss"], 2e-3 * float(ro["mr_cls"].abs().max()))
    state_close(rt, oracle, 1e-3, only=lambda n: "cls" in n)
    # ---- and back: a GAN checkpoint restores with its RMSProp slots, learning rate and step --------------------------
    gan_dir = str(tmp_path / "gan")
    os.makedirs(gan_dir)
    atr.dis_optimizer.set_lr(1.23e-4)
    atr.save(os.path.join(gan_dir, "model.cpkt"), gan_dir)
    slots_before = atr.dis_optimizer.slot_state()              # per-variable views (the arena's padding is not part of a checkpoint)
    assert any(float(np.abs(v - 1.0).max()) > 0 for k, v in slots_before.items() if k.endswith("/RMSProp"))
    atr.dis_optimizer.ms.fill_(1.0)
    atr.dis_optimizer.set_lr(3e-4)
    anet.restore(os.path.join(gan_dir, "latest.npz"))
    n = atr.load_optimizer_state(anet.last_restored)
    slots_after = atr.dis_optimizer.slot_state()
    assert n > 0 and all(np.array_equal(slots_before[k], slots_after[k]) for k in slots_before)
    assert abs(atr.dis_optimizer.get_lr() - 1.23e-4) < 1e-10
    atr.dis_optimizer.ms.fill_(1.0)
    anet.restore(os.path.join(gan_dir, "latest.npz"), clear_rms=True)          # 'RMS' names are filtered out (:541)
    assert atr.load_optimizer_state(anet.last_restored, clear_rms=True) == 0 and float(atr.dis_optimizer.ms.min()) == 1.0


def test_evaluation_path_matches_oracle():
    """Trainer.evaluate: CT slices through the DAM + shared back half in inference mode (every BN folded into its
    convolution's epilogue on the tcgen05 path), hard Dice with background (lib.py:96-110) and the confusion matrix"""
    import pnp_b200  # noqa: F401
    from pnp_b200 import runtime as rt, adversarial as adv, functional as F
    from pnp_b200.train_gan import configure
    from oracle.pnp_graphs import OracleAdversarial, init_numpy_params, synthetic_images, synthetic_labels
    from oracle.tf14_numpy import label_decomp
    from oracle import tf14_torch as T
    rt.set_conv_backend("auto")
    ws, bns = OracleAdversarial.layout()
    P = init_numpy_params(ws, bns, 0, 0.05)
    _bn_noise(P, bns, 6)
    ck, nc, tc = configure("train-gan")
    net = adv.Full_DRN(channels=3, n_class=5, batch_size=B, cost_kwargs=ck, network_config=nc)
    rt.load_state_dict(P)
    trainer = adv.Trainer(net, num_cls=5, batch_size=B, opt_kwargs={"learning_rate": 3e-4}, train_config=tc)
    oracle = OracleAdversarial(P, B, lambda_mask_loss=0.3, dis_sub_iter=1, gen_sub_iter=1)
    ct = synthetic_images(B, 4321, 0.3, 0.8)
    y = torch.from_numpy(label_decomp(5, synthetic_labels(B, 99)))
    F.PROFILE = []
    st = trainer.evaluate(ct.to(DEV), y.to(DEV))
    n_conv = len(F.PROFILE)
    F.PROFILE = None
    with torch.no_grad():
        ref = oracle.segment(ct, "ct", 1.0, False)["logits"]
        logits = net.segment(ct.to(DEV), "ct", 1.0, front_bn=False)["logits"]
    check("evaluation logits", logits, ref, 1e-3)
    d_ref, arr_ref = T.dice_eval(ref.argmax(3), y, 5)
    print("  dice_eval %.6f (oracle %.6f), %d convolution launches in one evaluation forward" % (st["dice_eval"], float(d_ref), n_conv))
    assert abs(st["dice_eval"] - float(d_ref)) <= 1e-3
    assert max(abs(a - float(b)) for a, b in zip(st["dice_arr"], arr_ref)) <= 1e-3
    pred, truth = ref.argmax(3).reshape(-1).numpy(), y.argmax(3).reshape(-1).numpy()
    cm_ref = np.zeros((5, 5), np.int64)
    np.add.at(cm_ref, (truth, pred), 1)
    cm = st["confusion_matrix"]
    assert cm.sum() == B * 256 * 256 and np.abs(cm - cm_ref).sum() <= 2e-3 * cm.sum(), (cm, cm_ref)
    rt.set_conv_backend("auto")


def test_per_volume_evaluation_matches_oracle():
    """Trainer.test_eval_volume (adversarial.py:993-1052 without the NIfTI reader): a synthetic [256,256,D] subject, frames fed with
    their neighbours as channels, confusion matrix over the subject -> per-class Dice / Jaccard (lib.py:121-152) == the same
    protocol run on the oracle"""
    import pnp_b200  # noqa: F401
    from pnp_b200 import runtime as rt, adversarial as adv
    from pnp_b200.data import label_maps
    from pnp_b200.lib import _dice, _jaccard
    from pnp_b200.train_gan import configure
    from oracle.pnp_graphs import OracleAdversarial, init_numpy_params
    rt.set_conv_backend("auto")
    ws, bns = OracleAdversarial.layout()
    P = init_numpy_params(ws, bns, 0, 0.05)
    _bn_noise(P, bns, 6)
    ck, nc, tc = configure("train-gan")
    net = adv.Full_DRN(channels=3, n_class=5, batch_size=B, cost_kwargs=ck, network_config=nc)
    rt.load_state_dict(P)
    trainer = adv.Trainer(net, num_cls=5, batch_size=B, opt_kwargs={"learning_rate": 3e-4}, train_config=tc)
    D = 7                                   # 5 usable frames -> floor(7 / 2) = 3 batches of 2, as the reference counts them
    rng = np.random.RandomState(3)
    raw = rng.randn(256, 256, D).astype(np.float32)
    raw_y = np.transpose(label_maps(D, 55), (1, 2, 0)).copy()
    dice, jac, cm, pred = trainer.test_eval_volume(raw, raw_y, flip_correction=True, shuffle_seed=9)
    # the same protocol on the oracle
    oracle = OracleAdversarial(P, B, lambda_mask_loss=0.3, dis_sub_iter=1, gen_sub_iter=1)
    r2, y2 = np.flip(np.flip(raw, 0), 1), np.flip(np.flip(raw_y, 0), 1)
    frames = list(range(1, D - 1))
    np.random.RandomState(9).shuffle(frames)
    cm_ref = np.zeros((5, 5), np.int64)
    for ii in range(D // B):
        idx = frames[ii * B:(ii + 1) * B]
        vol = np.zeros((B, 256, 256, 3), np.float32)
        sl = np.zeros((B, 256, 256), np.int64)
        for k, jj in enumerate(idx):
            vol[k] = r2[..., jj - 1:jj + 2]
            sl[k] = y2[..., jj]
        with torch.no_grad():
            p_ref = oracle.segment(torch.from_numpy(vol), "ct", 1.0, False)["logits"].argmax(3).numpy()
        np.add.at(cm_ref, (sl.reshape(-1), p_ref.reshape(-1)), 1)
    print("  per-class Dice   ours", np.round(dice, 5), "oracle", np.round(_dice(cm_ref), 5))
    print("  per-class Jaccard ours", np.round(jac, 5), "oracle", np.round(_jaccard(cm_ref), 5))
    assert cm.sum() == cm_ref.sum() == (D // B) * B * 256 * 256
    assert np.abs(cm - cm_ref).sum() <= 2e-3 * cm.sum()
    assert np.abs(dice - _dice(cm_ref)).max() <= 1e-3 and np.abs(jac - _jaccard(cm_ref)).max() <= 1e-3
