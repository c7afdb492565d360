"""CPU (-m "not gpu"): the C-ABI shared library builds, loads, and exports exactly the symbols include/adapter4rec.h
declares, and the ctypes prototype table covers all of them.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "adapter4rec.h")


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"A4R_API\s+[\w\s\*]+?\b(a4r_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from adapter4rec_b200 import _build
    return _build.build()


def test_header_declares_symbols():
    syms = header_symbols()
    assert len(syms) >= 20 and "a4r_gemm_bf16_tn" in syms and "a4r_score_topk" in syms


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT (a4r_\w+)", out)))
    assert exported == header_symbols()


def test_ctypes_table_matches_header(lib_path):
    from adapter4rec_b200 import lib
    assert sorted(lib.PROTOTYPES.keys()) == header_symbols()
    handle = lib.get_lib()
    assert handle.a4r_version() == 100
    assert handle.a4r_last_error_string() == b""
    assert handle.a4r_launch_count() == 0


def test_no_gpu_calls_fail_loudly(lib_path):
    """Without a CUDA device every compute entry point must return an error code, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from adapter4rec_b200 import lib
    handle = lib.get_lib()
    assert handle.a4r_device_check() != 0
    assert b"cuda" in handle.a4r_last_error_string().lower() or len(handle.a4r_last_error_string()) > 0
    g = lib.GemmArgs()
    assert handle.a4r_gemm_bf16_tn(ctypes.byref(g), None) == lib.A4R_EINVAL   # argument validation precedes launch


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under adapter4rec_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "adapter4rec_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "transrec_oracle" not in src and "oracle" + os.sep not in src, os.path.join(dirpath, f)


def test_missing_library_raises(monkeypatch, tmp_path):
    from adapter4rec_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        lib.get_lib()


def test_entry_script_flags_equal_the_reference():
    """adapter4rec_b200.parameters.parse_args([]) == Downstream/Text/parameters.py parse_args() (names and defaults);
    tests/golden/text_flags.json was dumped from the reference's own parser in the build container."""
    import json
    from adapter4rec_b200.parameters import parse_args
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "text_flags.json")))
    assert vars(parse_args([])) == ref


def test_cv_entry_script_flags_equal_the_reference():
    """adapter4rec_b200.cv.parameters.parse_args([]) == Downstream/CV/parameters.py parse_args() (names, types and defaults);
    tests/golden/cv_flags.json is the reference parser's own dump (tests/golden/make_cv_flags.py)."""
    import json
    from adapter4rec_b200.cv.parameters import parse_args
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "cv_flags.json")))
    got = vars(parse_args([]))
    assert got == ref
    assert {k: type(v) for k, v in got.items()} == {k: type(v) for k, v in ref.items()}


def test_cv_parameter_groups_follow_the_reference_rules():
    """run_adapter.py:487-510: names with 'image_net' go to the image group unless they are classifier-like ('fc' / 'classifier' /
    'decoder_pred' in the name); 'adapter' (and only 'adapter': LoRA factors are NOT adapters here) selects the adapter groups.
    A reference quirk the mirror keeps: the Houlsby bottleneck's parameters are called fc_down / fc_up, so the 'fc' test sends
    the ViT's adapters to the adapter_recsys group (adapter_sasrec_lr); only adapters without 'fc' in their names (Compacter's
    down_sampler / up_sampler) train with --adapter_cv_lr."""
    import torch
    from adapter4rec_b200.cv.run_adapter import group_parameters_cv

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.names = ["cv_encoder.image_net.vit.encoder.layer.0.output.adapter.fc_down.weight",
                          "cv_encoder.image_net.vit.encoder.layer.0.output.adapter.down_sampler.W_left",
                          "cv_encoder.image_net.vit.encoder.layer.0.attention.attention.query.lora_A",
                          "cv_encoder.image_net.classifier.weight", "user_encoder.transformer_encoder.layer_norm.weight",
                          "user_encoder.transformer_encoder.transformer_blocks.0.adapter1.fc_up.bias", "frozen.weight"]

        def named_parameters(self, *a, **k):
            for n in self.names:
                p = torch.nn.Parameter(torch.zeros(1))
                p.requires_grad = n != "frozen.weight"
                yield n, p

    g = {k: [n for n, _ in v] for k, v in group_parameters_cv(M()).items()}
    assert g["adapter_bert"] == ["cv_encoder.image_net.vit.encoder.layer.0.output.adapter.down_sampler.W_left"]
    assert g["bert"] == ["cv_encoder.image_net.vit.encoder.layer.0.attention.attention.query.lora_A"]
    assert g["recsys"] == ["cv_encoder.image_net.classifier.weight", "user_encoder.transformer_encoder.layer_norm.weight"]
    assert g["adapter_recsys"] == ["cv_encoder.image_net.vit.encoder.layer.0.output.adapter.fc_down.weight",
                                   "user_encoder.transformer_encoder.transformer_blocks.0.adapter1.fc_up.bias"]
