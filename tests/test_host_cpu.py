"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/icap.h declares
with matching argument lists, the drop-in module reproduces the reference's state_dict layout, and the
product path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import icap_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = icap_loader.load()
N = pkg._native

CTYPE = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "uint64_t": ctypes.c_uint64}


def _header_decls():
    h = open(os.path.join(ROOT, "include", "icap.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    decls = {}
    for ret, name, args in re.findall(r"^(int|const char\*) (icap_\w+)\((.*?)\);", h, flags=re.S | re.M):
        args = " ".join(args.split())
        decls[name] = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
    return decls


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    decls = _header_decls()
    assert len(decls) >= 22
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/icap.h but not exported by libicap.so"
    assert lib.icap_version() == 100


def test_ctypes_signatures_match_header():
    decls = _header_decls()
    for name, args in decls.items():
        if name == "icap_last_error":
            continue
        sig = N.SIGNATURES[name]
        assert len(sig) == len(args), name
        for a, ct in zip(args, sig):
            if "*" in a:
                assert ct is ctypes.c_void_p, (name, a)
            else:
                base = a.replace("const ", "").split()[0]
                assert ct is CTYPE[base], (name, a, ct)


def test_argument_errors_are_reported_without_a_gpu():
    # validation happens before any CUDA call, so this is safe on a CPU-only box
    with pytest.raises(N.IcapError) as e:
        N.call("icap_add_ln_fwd", 0, 0, 4, 30, 16, None, 1, 16, 16, None, 16, None, None, 0, 0.0, 0, None, 1e-6, None)
    assert "multiple of 4" in str(e.value)
    with pytest.raises(N.IcapError):
        N.call("icap_gemm", 1, 1, 1, 0, 8, 8, 16, 8, 16, 8, 16, 8, 0, None, 0, None, 0, 0, 1, None)


def test_state_dict_layout_matches_reference_golden():
    for name in ("tiny_default", "tiny_cfgpy", "tiny_variants"):
        g = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
        m = pkg.Transformer(device=torch.device("cpu"), **g["ctor"])
        assert list(m.state_dict().keys()) == g["state_dict_keys"]
        for k, v in m.state_dict().items():
            assert v.shape == g["state_dict"][k].shape, k
        m.load_state_dict(g["state_dict"])
        for k, v in m.state_dict().items():
            assert torch.equal(v, g["state_dict"][k]), k
        # all parameters are views of one flat buffer, q/k/v weights adjacent (packed QKV GEMM)
        base = m._flat.data_ptr()
        for k, q in m.named_parameters():
            assert base <= q.data_ptr() < base + m._flat.numel() * 4
        for k in m._offsets:
            if k.endswith("q_linear.weight"):
                kk, vv = k.replace("q_linear", "k_linear"), k.replace("q_linear", "v_linear")
                assert m._offsets[kk] == m._offsets[k] + m._shapes[k][0] * m._shapes[k][1]
                assert m._offsets[vv] == m._offsets[kk] + m._shapes[kk][0] * m._shapes[kk][1]


def test_init_statistics_follow_reference():
    torch.manual_seed(0)
    m = pkg.Transformer(num_vocab=1000, max_length=22, encode_dim_positions=84, encode_dim_features=2048,
                        device=torch.device("cpu"), output_name="x")
    sd = m.state_dict()
    assert abs(float(sd["encoder.encoder.0.multihead_attention.q_linear.weight"].std()) - (2 / 1024) ** 0.5) < 2e-3
    assert abs(float(sd["classifer.weight"].std()) - (2 / 1512) ** 0.5) < 2e-3
    assert float(sd["decoder.word_embedding.weight"][0].abs().sum()) == 0
    assert abs(float(sd["decoder.word_embedding.weight"][1:].std()) - 1) < 2e-2
    assert float(sd["encoder.feature_embedding.weight"].abs().max()) <= 1 / 2048 ** 0.5 + 1e-7
    assert torch.equal(sd["encoder.norm.weight"], torch.ones(512))
    assert sum(p.numel() for p in m.parameters()) == 1000 * 512 * 2 + 1000 + 55_707_408 - (10000 * 512 * 2 + 10000)


def test_no_cpu_fallback():
    g = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_default.pt"), weights_only=False)
    m = pkg.Transformer(device=torch.device("cpu"), **g["ctor"])
    with pytest.raises(N.IcapError):
        m(g["features"], g["positions"], g["captions"])
    with pytest.raises(N.IcapError):
        m.generate_caption_vector(g["features"], g["positions"])
    with pytest.raises(N.IcapError):
        m.beam_search(g["features"], g["positions"], beam_size=3)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "image-caption_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, fn)


def test_resnet_extractor_state_dict_matches_torchvision():
    """the drop-in ResnetExtractor exposes exactly the parameter / buffer names and shapes of the
    reference's nn.Sequential(*resnet101.children()[:9]) (preprocess.py:37-40), so pretrained trunks load unchanged."""
    import sys
    tv = pytest.importorskip("torchvision")
    pkg_dir = os.path.join(ROOT, "image-caption_b200")
    sys.path.insert(0, pkg_dir)
    try:
        from core.preprocess import ResnetExtractor
        ext = ResnetExtractor()
        ref = torch.nn.Sequential(*list(tv.models.resnet101(weights=None).children())[:9])
        ours, theirs = ext.submodule.state_dict(), ref.state_dict()
        assert list(ours.keys()) == list(theirs.keys())
        for k in ours:
            assert tuple(ours[k].shape) == tuple(theirs[k].shape), k
        ext.submodule.load_state_dict(theirs)
        assert ext.image_size == 224 and ext.training
    finally:
        sys.path.remove(pkg_dir)
        for m in [k for k in sys.modules if k == "core" or k.startswith("core.")]:
            sys.modules.pop(m)
