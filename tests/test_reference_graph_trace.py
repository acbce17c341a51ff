"""Architecture parity against the reference's OWN graph-construction code.

tests/golden/reference_graph_trace.json is the canonical layer list obtained by executing /root/reference/adversarial.py
(+ layers.py, ops.py) unmodified under a recording shim of tensorflow (tests/golden/make_reference_graph_trace.py).  Here the
PRODUCT's graph code (pnp_b200.adversarial / networks / layers / ops) runs on CPU with recording stand-ins for its six kernel
entry points (functional.conv_layer, res_block, max_pool2, phase_shift, disc_input, fc) on meta tensors, and the two lists must
agree event by event: filter variable name + shape, stride, dilation, padding rule, dropout keep_prob source, batch-norm scope /
training switch / trainable flag, skip kind, activation, pooling, PS parameters, discriminator input channel layout, matmuls,
plus the full variable table (376 names, shapes, trainable flags) and the weight lists behind the L2 terms.
Nothing here needs a GPU or reads /root/reference."""
import json
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_graph_trace.json")) as _f:
    REF = json.load(_f)

CFG = {"mr_front_trainable": False, "ct_front_trainable": True, "joint_trainable": False, "cls_trainable": True, "m_cls_trainable": True}
COST = {"regularizer": 1e-4, "gan_regularizer": 1e-4, "miu_dis": 1e-3, "miu_gen": 2e-3, "lambda_mask_loss": 0.1}
KEEP_PH = 0.4375          # stands for the keep_prob placeholder (a value no Python default of the reference uses)
B = REF["batch"]


class _Tracer(object):
    """recording stand-ins for the kernels behind pnp_b200.functional"""

    def __init__(self, F, rt):
        self.F, self.rt = F, rt
        self.events = []
        self.names = {id(v): k for k, v in rt.graph.vars.items()}
        self.src = {}                      # id(tensor) -> provenance label

    def name(self, t):
        return self.names[id(t)]

    def _conv(self, x, W, cfg, skip):
        F = self.F
        _, g = F._geometry(tuple(x.shape), tuple(W.shape), cfg)
        bn = cfg.bn
        ev = {"op": "conv", "w": self.name(W), "wshape": list(W.shape), "w_trainable": bool(W.requires_grad), "stride": cfg.stride, "dil": cfg.dil,
              "padding": cfg.padding, "in": list(x.shape[1:]), "out": [g.Ho, g.Wo, g.Cout], "keep": float(cfg.keep_prob),
              "bn": None if bn is None else self.name(bn.gamma).rsplit("/", 1)[0],
              "bn_train": None if bn is None else bool(cfg.bn_training), "bn_trainable": None if bn is None else bool(bn.gamma.requires_grad),
              "act": {F.ACT_NONE: "none", F.ACT_RELU: "relu", F.ACT_LRELU: "lrelu0.2"}[cfg.act], "skip": skip,
              "input_layout": getattr(x, "_layout", None)}
        self.events.append(ev)
        y = torch.empty(x.shape[0], g.Ho, g.Wo, g.Cout, device="meta")
        self.src[id(y)] = "out_of:" + ev["w"]
        self._keep = getattr(self, "_keep", []) + [y]          # keep ids alive / unique
        return y

    def conv_layer(self, x, W, cfg, skip=None):
        assert skip is None
        return self._conv(x, W, cfg, "none")

    def res_block(self, x, W1, W2, cfg1, cfg2):
        assert cfg1.skip_off == 0
        h = self._conv(x, W1, cfg1, "none")
        return self._conv(h, W2, cfg2, ("pad%d" % cfg2.skip_off) if cfg2.skip_off else "identity")

    def max_pool2(self, x):
        self.events.append({"op": "maxpool", "k": 2, "stride": 2, "in": list(x.shape[1:]), "out": [x.shape[1] // 2, x.shape[2] // 2, x.shape[3]]})
        y = torch.empty(x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3], device="meta")
        self._keep = getattr(self, "_keep", []) + [y]
        return y

    def _ps_event(self, X, r, n_channel):
        assert X.shape[3] == r * r * n_channel
        self.events.append({"op": "PS", "r": r, "n_channel": n_channel, "in": list(X.shape[1:]), "out": [X.shape[1] * r, X.shape[2] * r, n_channel],
                            "of": self.src.get(id(X))})

    def phase_shift(self, X, r, n_channel, batch_size):
        assert batch_size == B
        self._ps_event(X, r, n_channel)
        y = torch.empty(X.shape[0], X.shape[1] * r, X.shape[2] * r, n_channel, device="meta")
        self.src[id(y)] = "PS(%s)" % self.src.get(id(X))
        self._keep = getattr(self, "_keep", []) + [y]
        return y

    def disc_input(self, c4, c6, b7, c9, logits, batch_size, r=8):
        assert batch_size == B and r == 8
        layout = []
        for t, ntile in zip((c4, c6, b7, c9), (3, 1, 1, 1)):       # the fused gather of functional._DiscInputFn
            g = t.shape[3] // (r * r)
            self._ps_event(t, r, g)
            layout += [["PS(%s)" % self.src.get(id(t)), g]] * ntile
        layout += [[self.src.get(id(logits)), logits.shape[3]], ["argmax(%s)" % self.src.get(id(logits)), 1]]
        y = torch.empty(logits.shape[0], logits.shape[1], logits.shape[2], sum(c for _, c in layout), device="meta")
        y._layout = layout
        self._keep = getattr(self, "_keep", []) + [y]
        return y

    def fc(self, x, w):
        self.events.append({"op": "fc", "w": self.name(w), "wshape": list(w.shape), "w_trainable": bool(w.requires_grad), "in": list(x.shape[1:])})
        return torch.empty(x.shape[0], w.shape[1], device="meta")


@pytest.fixture(scope="module")
def product():
    from pnp_b200 import adversarial as A, functional as F, runtime as rt
    net = A.Full_DRN(3, 5, B, cost_kwargs=dict(COST), network_config=dict(CFG))
    tr = _Tracer(F, rt)
    saved = {k: getattr(F, k) for k in ("conv_layer", "res_block", "max_pool2", "phase_shift", "disc_input", "fc")}
    for k in saved:
        setattr(F, k, getattr(tr, k))
    try:
        out = {}

        def run(label, fn):
            tr.events = []
            r = fn()
            out[label] = tr.events
            return r
        x = torch.empty(B, 256, 256, 3, device="meta")
        # distinct BN switch values per sub-graph prove which switch feeds which layers
        mr = run("mr", lambda: net.segment(x, "mr", KEEP_PH, front_bn=True, joint_bn=False))
        ct = run("ct", lambda: net.segment(x, "ct", KEEP_PH, front_bn=False, joint_bn=True))
        run("cls", lambda: net.create_classifier(ct["c4_2"], ct["c6_2"], ct["b7"], ct["c9_2"], ct["logits"]))
        run("mask", lambda: net.create_mask_critic(ct["logits"]))
    finally:
        for k, v in saved.items():
            setattr(F, k, v)
    return net, rt, out


def _ref_events(section):
    return [e for e in REF["events"] if e["section"] == section]


def _norm_ref(ev, bn_switch):
    """reference event -> the product's vocabulary"""
    e = dict(ev)
    if e["op"] == "conv":
        e["keep"] = KEEP_PH if e["keep"] == "ph:keep_prob" else e["keep"]
        if isinstance(e["bn_train"], str):
            e["bn_train"] = bn_switch[e["bn_train"]]
    return e


CONV_KEYS = ("w", "wshape", "w_trainable", "stride", "dil", "padding", "in", "out", "keep", "bn", "bn_train", "bn_trainable", "act", "skip")


def _compare(ref_list, got_list, bn_switch, what):
    assert len(ref_list) == len(got_list), "%s: %d reference layers vs %d here" % (what, len(ref_list), len(got_list))
    for i, (r, g) in enumerate(zip(ref_list, got_list)):
        r = _norm_ref(r, bn_switch)
        assert r["op"] == g["op"], (what, i, r["op"], g["op"])
        if r["op"] == "conv":
            for k in CONV_KEYS:
                assert r[k] == g[k], "%s layer %d (%s): %s reference %r vs %r" % (what, i, r["w"], k, r[k], g[k])
            assert r["bn_decay"] in (None, 0.9)
            if r.get("input_layout"):
                assert g["input_layout"] == r["input_layout"], (what, i, r["input_layout"], g["input_layout"])
        elif r["op"] == "maxpool":
            assert (r["k"], r["stride"], r["in"], r["out"]) == (g["k"], g["stride"], g["in"], g["out"]), (what, i, r, g)
            assert r["padding"] == "SAME"          # even extents: SAME == VALID for the 2x2/2 pool
        elif r["op"] == "PS":
            for k in ("r", "n_channel", "in", "out", "of"):
                assert r[k] == g[k], (what, i, k, r[k], g[k])
            assert r["batch_size_arg"] == "self.batch_size"
        elif r["op"] == "fc":
            for k in ("w", "wshape", "w_trainable", "in"):
                assert r[k] == g[k], (what, i, k, r[k], g[k])


def test_reference_init_fails_where_the_survey_says():
    assert "predicter" in REF["init_error_after_classifier"]           # adversarial.py:102


def test_mr_and_ct_front_halves_match_the_reference_zip_network(product):
    _, _, got = product
    zipn = _ref_events("create_zip_network#1")
    assert len(zipn) == 48
    mr_ref, ct_ref = zipn[:24], zipn[24:]
    assert mr_ref[0]["input_src"] == "ph:mr_ph" and ct_ref[0]["input_src"].startswith("ph:")
    _compare(mr_ref, got["mr"][:24], {"ph:main_batchnorm_training_switch": True}, "MR front (groups 1-6)")
    _compare(ct_ref, got["ct"][:24], {"ph:adapt_batchnorm_training_switch": False}, "CT front (adapt 1-6)")


def test_second_half_matches_both_reference_calls(product):
    _, _, got = product
    h1, h2 = _ref_events("create_second_half#1"), _ref_events("create_second_half#2")
    strip = lambda evs: [{k: v for k, v in e.items() if k not in ("section", "of")} for e in evs]
    assert strip(h1) == strip(h2)                                        # AUTO_REUSE: the very same layers on the other stream
    _compare(h1, got["ct"][24:], {"ph:joint_batchnorm_training_switch": True}, "second half (CT stream)")
    _compare(h2, got["mr"][24:], {"ph:joint_batchnorm_training_switch": False}, "second half (MR stream)")


def test_feature_discriminator_matches_the_reference(product):
    _, _, got = product
    c1, c2 = _ref_events("create_classifier#1"), _ref_events("create_classifier#2")
    assert [e["w"] for e in c1 if e["op"] != "PS"] == [e["w"] for e in c2 if e["op"] != "PS"]
    _compare(c1, got["cls"], {}, "feature discriminator")
    first = [e for e in got["cls"] if e["op"] == "conv"][0]
    assert sum(c for _, c in first["input_layout"]) == 32 == first["in"][2]


def test_mask_critic_matches_the_reference(product):
    _, _, got = product
    m1, m2 = _ref_events("create_mask_critic#1"), _ref_events("create_mask_critic#2")
    assert [e.get("w") for e in m1] == [e.get("w") for e in m2]
    _compare(m1, got["mask"], {}, "mask critic")


def test_variable_table_and_weight_lists_match_the_reference(product):
    net, rt, _ = product
    ref_vars = REF["variables"]
    # same 376 variables.  Global creation order differs harmlessly (the product builds the shared second half before the CT
    # front and a stage's filters before its batch-norm variables); nothing on the path depends on it: savers and the optimizer
    # variable lists go by name / scope substring (adversarial.py:478-501)
    ref_names = [v["name"] for v in ref_vars]
    assert sorted(ref_names) == sorted(rt.graph.order) and len(set(ref_names)) == len(ref_names) == 376
    top = lambda n: n.split("/")[0]
    is_w = lambda n: n.rsplit("/", 1)[1] not in ("beta", "gamma", "moving_mean", "moving_variance")
    for scope in sorted(set(top(n) for n in ref_names)):       # filter weights keep their creation order inside a scope
        assert [n for n in ref_names if top(n) == scope and is_w(n)] == [n for n in rt.graph.order if top(n) == scope and is_w(n)], scope
    for v in ref_vars:
        t = rt.graph.vars[v["name"]]
        assert list(t.shape) == v["shape"], v["name"]
        assert bool(t.requires_grad) == v["trainable"], v["name"]
    names = {id(v): k for k, v in rt.graph.vars.items()}
    for key in ("mr_front_weights", "ct_front_weights", "cls_weights", "m_cls_weights", "joint_weights"):
        assert [names[id(w)] for w in getattr(net, key)] == REF["weight_lists"][key], key
    # the reference's initialisers: weight_variable (stddev 0.01, tf.Variable) for MR groups 1-4, sharable (0.1) elsewhere
    kinds = {v["name"]: (v["kind"], v["stddev"]) for v in ref_vars if v["kind"] != "batch_norm"}
    assert all(kinds[n] == ("tf.Variable", 0.01) for n in kinds if n.split("/")[0] in ("group_1", "group_2", "group_3", "group_4"))
    assert all(kinds[n] == ("tf.get_variable", 0.1) for n in kinds if n.split("/")[0] not in ("group_1", "group_2", "group_3", "group_4"))


# ------------------------------------------------------------------------------------------------
# source segmenter (source_segmenter.py:48-273; the file's head up to `class Trainer` executed verbatim)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def product_segmenter():
    from pnp_b200 import source_segmenter as S, functional as F, runtime as rt
    args = REF["source_segmenter"]["ctor_args"]
    net = S.Full_DRN(3, 5, B, cost_kwargs={"cross_flag": True, "miu_cross": 1.0, "dice_flag": True, "miu_dice": 1.0, "regularizer": 1e-4}, **args)
    tr = _Tracer(F, rt)
    saved = {k: getattr(F, k) for k in ("conv_layer", "res_block", "max_pool2", "phase_shift", "disc_input", "fc")}
    for k in saved:
        setattr(F, k, getattr(tr, k))
    try:
        net.forward(torch.empty(B, 256, 256, 3, device="meta"), keep_prob=KEEP_PH, main_bn=True, adapt_bn=False)
    finally:
        for k, v in saved.items():
            setattr(F, k, v)
    return net, rt, tr.events


def test_source_segmenter_file_state_is_what_the_survey_says():
    s = REF["source_segmenter"]
    assert s["syntax_error"]["line"] == 611 and s["executed_lines"] == 302


def test_source_segmenter_network_matches_the_reference(product_segmenter):
    _, _, got = product_segmenter
    ref = REF["source_segmenter"]["events"]
    assert len(ref) == 37
    _compare(ref, got, {"ph:adapt_batchnorm_training_switch": False, "ph:main_batchnorm_training_switch": True}, "source segmenter")
    # groups 1-4 follow adapt_trainable / adapt_bn, groups 5+ main_trainable / main_bn
    for e in ref:
        if e["op"] == "conv":
            early = e["w"].split("/")[0] in ("group_1", "group_2", "group_3", "group_4")
            assert e["w_trainable"] == early, e["w"]
            if e["bn"]:
                assert e["bn_train"] == ("ph:adapt_batchnorm_training_switch" if early else "ph:main_batchnorm_training_switch"), e["w"]


def test_source_segmenter_variables_and_l2_list_match_the_reference(product_segmenter):
    net, rt, _ = product_segmenter
    ref = REF["source_segmenter"]
    names = [v["name"] for v in ref["variables"]]
    assert sorted(names) == sorted(rt.graph.order) and len(names) == 153
    for v in ref["variables"]:
        t = rt.graph.vars[v["name"]]
        assert list(t.shape) == v["shape"] and bool(t.requires_grad) == v["trainable"], v["name"]
        if v["kind"] == "tf.Variable":
            assert v["stddev"] == 0.01                           # weight_variable, layers.py:46-48
    # anonymous batch-norm scopes number in creation order: BatchNorm, BatchNorm_1, ... BatchNorm_29 at the top level
    bn_scopes = [n.rsplit("/", 1)[0] for n in names if n.endswith("/beta")]
    assert bn_scopes == ["BatchNorm"] + ["BatchNorm_%d" % i for i in range(1, 30)]
    # the L2 list with the reference's quirk (source_segmenter.py:132-135): wr4_4 twice, wr4_3 never
    ids = {id(v): k for k, v in rt.graph.vars.items()}
    assert [ids[id(w)] for w in net.conv_weights] == ref["conv_weights"]
    assert ref["conv_weights"].count("group_4/Variable_3") == 2 and "group_4/Variable_2" not in ref["conv_weights"]
