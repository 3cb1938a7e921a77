"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every
symbol include/slcl.h declares, the ctypes table mirrors the header, the Python
package keeps the reference's names/signatures, and the product path fails
loudly without a GPU (no CPU fallback)."""
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "slcl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(slcl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from slcl import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libslcl.so lacks {name}"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    assert lib.slcl_version() == 115
    assert lib.slcl_strerror(0) == b"ok"
    assert b"invalid" in lib.slcl_strerror(-1)
    assert lib.slcl_proto_workspace_bytes(1 << 20) >= (1 << 20) // 256 * 16
    assert lib.slcl_proto_workspace_bytes(0) == 0
    assert lib.slcl_compact_workspace_bytes(10000, 4) > 0


def test_ctypes_table_matches_the_header_prototypes():
    """Every prototype of include/slcl.h against slcl/_lib.py: the same number of parameters, pointers where the header
    has pointers, 64-bit integers / size_t / float / double / int where it has those (an ABI drift would otherwise only
    show up as a crash on the GPU box)."""
    import ctypes as C
    from slcl import _lib
    txt = open(os.path.join(ROOT, "include", "slcl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = []
    for name in _lib.SIGNATURES:
        m = re.search(r"\b" + name + r"\s*\(([^;{]*?)\)\s*;", txt, flags=re.S)
        assert m, f"no prototype of {name} in include/slcl.h"
        protos.append((name, m.group(1)))

    def kind(decl):
        decl = " ".join(decl.split())
        if "*" in decl or decl.startswith("slcl_stream_t"):
            return "ptr"
        for key, name in (("int64_t", "i64"), ("size_t", "size"), ("double", "f64"), ("float", "f32"), ("int", "i32")):
            if re.search(r"\b" + key + r"\b", decl):
                return name
        raise AssertionError(f"unparsed parameter {decl!r}")

    ckind = {C.c_void_p: "ptr", C.c_char_p: "ptr", C.c_int64: "i64", C.c_size_t: "size", C.c_double: "f64", C.c_float: "f32",
             C.c_int: "i32"}
    for name, params in protos:
        params = params.strip()
        want = [] if params in ("", "void") else [kind(p) for p in params.split(",")]
        restype, argtypes = _lib.SIGNATURES[name]
        got = []
        for a in argtypes:
            if a in ckind:
                got.append(ckind[a])
            else:          # POINTER(struct) and friends
                got.append("ptr")
        assert got == want, f"{name}: header {want} vs ctypes {got}"


def test_argument_validation_without_gpu():
    """Validation happens before any launch, so it can be exercised on a GPU-less host."""
    from slcl import _lib
    lib = _lib.load()
    assert lib.slcl_proto_fwd(None, None, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.slcl_class_sums(None, 1, 1, 1, None, None, 0, 0.0, None, 1, 4, None, None, 0, None) == -1
    assert lib.slcl_centroid_loss(None, None, 4, 32, 0, 1, 4, 1, None, None, None, None) == -1
    assert lib.slcl_compact_by_class(None, 10, 4, None, None, None, None, 0, None) == -1
    # pixel<->pixel entry points: null operands are rejected before anything touches the device; the size queries work
    assert lib.slcl_p2p_fwd(None, None, 128, 128, 64, None, None, None, 0, 1, None, None, 0.7, None, None, None, None, 0, None) == -1
    assert lib.slcl_p2p_bwd(None, None, 128, 128, 64, 64, None, None, None, None, 0, 1, None, None, 0.7, None, None, None, None,
                            None, None, 0, None) == -1
    assert lib.slcl_p2p_state_bytes(4096, 256) >= 4096 * 256 * 4
    assert lib.slcl_p2p_state_bytes(0, 256) == 0 and lib.slcl_p2p_workspace_bytes(4096, 16384, 512) == 0
    # peer-memory exchange: mailbox = 8 header words + 2 parities x world senders x capacity payload words; 1..16 ranks
    assert lib.slcl_peer_mailbox_bytes(8, 2) == (8 + 2 * 8 * 2) * 8 and lib.slcl_peer_mailbox_bytes(0, 2) == 0
    assert lib.slcl_peer_mailbox_bytes(17, 2) == 0 and lib.slcl_peer_mailbox_bytes(2, 1) == 0
    import ctypes as C
    assert lib.slcl_proto_rescale_peer(None, 1, None, None) == -1
    bad_rank = _lib.PeerT(16, 2, 2, 64, 1.0)
    assert lib.slcl_proto_rescale_peer(16, 1, C.byref(bad_rank), None) == -1       # rank out of range
    small = _lib.PeerT(16, 0, 2, 8, 1.0)
    assert lib.slcl_peer_allreduce_f64(16, 5, C.byref(small), None) == -1          # 2*n > capacity_words
    assert lib.slcl_class_centres_update(None, 1, 1, 1, None, 4, None, 0.9, None, None, None, None, 0, None) == -1
    assert lib.slcl_centroids_fwd(None, 1, 1, 1, None, None, 0, 0.0, None, 1, 4, None, 0.9, None, None, None, None, None, 0,
                                  None) == -1
    assert lib.slcl_p2p_workspace_bytes(4096, 16384, 256) > 0
    # sampler bookkeeping / row plumbing added in round 2
    assert lib.slcl_sample_balanced_workspace_bytes(65536, 5) > 0 and lib.slcl_sample_balanced_workspace_bytes(65536, 0) == 0
    assert lib.slcl_sample_balanced(None, None, 100, 4, 8, None, None, 8, None, None, None, 0, None) == -1
    assert lib.slcl_sample_balanced(16, 16, 100, 4, 0, 16, 16, 8, 16, 16, 16, 1 << 20, None) == -1      # per_a must be >= 1
    assert lib.slcl_sample_balanced(16, 16, 100, 4, 8, 16, 16, 8, None, None, 16, 1 << 20, None) == -1   # quota b without outputs
    assert lib.slcl_sample_balanced(16, 16, 100, 4, 8, 16, 16, 8, 16, 16, 16, 64, None) == -3            # workspace too small
    assert lib.slcl_self_maps(None, 10, None, 10, 100, None, None, None, None) == -1
    assert lib.slcl_rows_meta(None, 100, None, 10, None, None) == -1
    assert lib.slcl_scatter_rows_by_map(None, 1, 8, 16, 1, None, None, None, None, None, None, None, None) == -1
    assert lib.slcl_scatter_rows_by_map(16, 1, 8, 16, 1, 16, 16, 16, 16, None, None, 16, None) == -1      # half of set b
    assert lib.slcl_p2p_shift(None, 10, None, 10, 0.7, None, None) == -1
    assert lib.slcl_p2p_shift(16, 10, 16, 10, 0.0, 16, None) == -1                                        # temperature > 0


def test_reference_signatures_are_kept():
    from slcl import loss, losses, utils_
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(loss.MPCL.__init__) == ["self", "device", "num_class", "temperature", "m", "base_temperature", "easy_margin"]
    assert sig(loss.MPCL.forward) == ["self", "features", "labels", "class_center_feas", "pixel_sel_loc", "mask"]
    assert sig(loss.mpcl_loss_calc)[:6] == ["feas", "labels", "class_center_feas", "loss_func", "pixel_sel_loc", "tag"]
    assert sig(loss.ContrastiveLoss.__init__) == ["self", "tau", "n_class", "bg", "norm"]
    assert sig(loss.ContrastiveLoss.forward) == ["self", "centroid_s", "centroid_t", "bg", "split"]
    # the pixel<->pixel modules keep the reference's positional parameters; `n_class` is a trailing keyword extension
    assert sig(loss.SupConLoss.__init__)[:4] == ["self", "temperature", "contrast_mode", "base_temperature"]
    assert sig(loss.SupConLoss.forward) == ["self", "features", "labels"]
    assert sig(loss.LocalConLoss.__init__)[:3] == ["self", "temperature", "stride"]
    assert sig(loss.BlockConLoss.__init__)[:3] == ["self", "temperature", "block_size"]
    assert losses.SupConLoss is loss.SupConLoss
    assert sig(utils_.cal_centroid)[:15] == ["decoder_ft", "label", "previous_centroid", "momentum", "pseudo_label", "n_class",
                                             "partition", "threshold", "thd_w", "weighted_ave", "epoch", "max_epoch",
                                             "low_thd", "high_thd", "stdmin"]
    assert sig(utils_.update_class_center_iter)[:5] == ["cla_src_feas", "batch_src_labels", "class_center_feas", "m", "num_class"]
    assert sig(utils_.generate_pseudo_label) == ["cla_feas_trg", "class_centers", "pixel_sel_th"]
    m = loss.MPCL("cuda", num_class=4, temperature=.1, base_temperature=1, m=.4)
    assert abs(m.th - (-0.9210609940028851)) < 1e-12 and abs(m.mm - 0.15576733692346023) < 1e-12
    d = inspect.signature(utils_.cal_centroid).parameters
    assert d["momentum"].default == 0.95 and d["n_class"].default == 4 and d["partition"].default == 1
    assert inspect.signature(utils_.update_class_center_iter).parameters["m"].default == .2


def test_product_path_has_no_cpu_fallback():
    from slcl import loss, utils_
    from slcl._lib import SlclError
    feas = torch.randn(1, 8, 4, 4)
    lab = torch.zeros(1, 4, 4, dtype=torch.long)
    cc = torch.randn(4, 8)
    with pytest.raises((SlclError, NotImplementedError, RuntimeError)):
        loss.mpcl_loss_calc(feas, lab, cc, loss.MPCL("cpu", num_class=4), tag="source")
    with pytest.raises((SlclError, NotImplementedError, RuntimeError)):
        utils_.update_class_center_iter(feas, lab, cc)
    with pytest.raises((SlclError, NotImplementedError, RuntimeError)):
        utils_.generate_pseudo_label(feas, cc)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "soft-labeled-contrastive-learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|oracle[./]", txt, re.M), \
                    f"{f} references the oracle package"


def test_mpcl_error_behaviour_matches_reference():
    from slcl.loss import MPCL
    m = MPCL("cuda", num_class=4)
    with pytest.raises(ValueError):
        m(torch.randn(4, 8), torch.zeros(4), torch.randn(8, 4))
    with pytest.raises(ValueError):
        m(torch.randn(4, 1, 8), torch.zeros(4), torch.randn(8, 4), mask=torch.ones(4, 4))
    with pytest.raises(ValueError):
        m(torch.randn(4, 1, 8), torch.zeros(5), torch.randn(8, 4))
    with pytest.raises(RuntimeError):          # n_views > 1: the reference fails in the broadcast at utils/loss.py:552
        m(torch.randn(4, 2, 8), torch.zeros(4), torch.randn(8, 4))


def test_class_centre_state_layout(tmp_path):
    """a-9: NPY v1.0, '<f4', C order, (K, C); 128-byte header + 512-byte payload for (4, 32)."""
    from slcl import state
    src = os.path.join(ROOT, "tests", "golden", "class_center_ct_f0.npy")
    cc = state.load_class_centers(src, device="cpu")
    assert cc.shape == (4, 32) and cc.dtype == torch.float32
    norms = cc.norm(dim=1).tolist()
    assert [round(v, 2) for v in norms] == [2.98, 6.47, 6.53, 6.84]
    out = tmp_path / state.class_center_filename("/data/mmwhs", 0)
    assert out.name == "class_center_ct_f0.npy"
    assert state.class_center_filename("/x/mscmrseg/y", 3) == "class_center_bssfp_f3.npy"
    state.save_class_centers(str(out), cc)
    raw = open(out, "rb").read()
    assert len(raw) == 640 and raw[:8] == b"\x93NUMPY\x01\x00"
    assert raw == open(src, "rb").read()          # byte-identical round trip
    assert np.array_equal(np.load(out), cc.numpy())


def test_shard_range_covers_batch():
    from slcl.distributed import shard_range
    for n in (1, 7, 128, 129):
        for ws in (1, 2, 4, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_centre_state_in_checkpoint_dict(tmp_path):
    """f-3: the [K,C] state rides in the reference-style checkpoint dict and falls back to the .npy."""
    from slcl import state
    src = os.path.join(ROOT, "tests", "golden", "class_center_ct_f0.npy")
    cc = state.load_class_centers(src, device="cpu")
    ckpt = {"epoch": 3, "model_state_dict": {}, "optimizer_state_dict": {}}
    state.add_to_checkpoint(ckpt, cc * 2)
    path = tmp_path / "last.pt"
    torch.save(ckpt, path)
    back = torch.load(path)
    assert torch.equal(state.from_checkpoint(back, device="cpu"), cc * 2)
    old_style = {"epoch": 3, "model_state_dict": {}, "optimizer_state_dict": {}}
    assert torch.equal(state.from_checkpoint(old_style, device="cpu", fallback_npy=src), cc)
    with pytest.raises(KeyError):
        state.from_checkpoint(old_style, device="cpu")


def test_peer_mailbox_needs_a_process_group_and_the_exchange_is_a_noop_without_a_group():
    """slcl.peer / the loss-pair exchange on a host without GPUs: no silent degradation -- a mailbox cannot be built without
    an initialised process group, and group=None leaves the scalars alone (single-process semantics)."""
    from slcl.peer import PeerMailbox
    from slcl.functional import _exchange_loss_pair
    with pytest.raises(RuntimeError):
        PeerMailbox("cpu")
    scal = torch.tensor([1.0, 2.0, 3.0, 4.0])
    _exchange_loss_pair(scal, True, None)
    assert scal.tolist() == [1.0, 2.0, 3.0, 4.0]
