"""CPU-side checks (no GPU): the C-ABI library loads, exports every symbol include/ssv_b200.h declares, reports
sane sizes, refuses to compute without a B200 (no CPU fallback), and the python drop-in surface matches the
reference's signatures."""
import ctypes
import inspect
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "self-supervised-vision_b200", "ssv_b200", "libssv_b200.so")


@pytest.fixture(scope="module")
def cabi():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "self-supervised-vision_b200"), "-j8"], check=True)
    from ssv_b200 import _cabi
    return _cabi


def test_every_declared_symbol_is_exported(cabi):
    protos = cabi.parse_header()
    assert len(protos) >= 40
    raw = ctypes.CDLL(LIB)
    missing = [n for n in protos if not hasattr(raw, n)]
    assert not missing, f"declared in include/ssv_b200.h but not exported: {missing}"
    # and nothing exported with the public prefix is undeclared
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("ssvb_")}
    assert exported == set(protos), exported ^ set(protos)


def test_version_and_error_strings(cabi):
    L = cabi.lib()
    assert L.ssvb_version() == 110
    assert cabi.strerror(0) == "ok"
    for rc in (-1, -2, -3, -4, -5, -6):
        assert "ssv_b200" in cabi.strerror(rc)


def test_size_queries(cabi):
    L = cabi.lib()
    assert L.ssvb_ntxent_dpad(128) == 128 and L.ssvb_ntxent_dpad(96) == 128 and L.ssvb_ntxent_dpad(32) == 64
    assert L.ssvb_ntxent_mpad(32768) == 65536 and L.ssvb_ntxent_mpad(100) == 256
    # saved blob of BASELINE's NT-Xent config: bf16 zhat (16 MiB) + two M-vectors
    assert 16 * 2 ** 20 <= L.ssvb_ntxent_saved_bytes(32768, 128) <= 18 * 2 ** 20
    assert L.ssvb_ntxent_workspace_bytes(32768, 128) > 65536 * 128 * 4
    assert L.ssvb_ntxent_saved_bytes(0, 128) == 0
    assert L.ssvb_moco_workspace_bytes(256, 65536, 128) >= 65536 * 128 * 2
    assert L.ssvb_barlow_saved_bytes(2048, 8192) >= 8192 * 8192 * 2 + 2 * 2048 * 8192 * 2
    assert L.ssvb_sinkhorn_workspace_bytes(4096, 3000) > 0
    assert L.ssvb_swav_saved_bytes(512, 3000, 3000, 128) > 0
    # the distributed workspace shrinks with the number of local rows
    assert L.ssvb_ntxent_dist_workspace_bytes(8, 4096, 128) < L.ssvb_ntxent_dist_workspace_bytes(1, 32768, 128)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(cabi):
    import ssv_b200
    L = cabi.lib()
    assert L.ssvb_device_check() == -5  # SSVB_ERR_ARCH
    # compute entry points refuse before touching any pointer
    assert L.ssvb_rowdot_fwd(0, None, None, 4, 4, 4, 4, None, None, 0, None) == -5
    assert L.ssvb_ntxent_fwd(None, None, 8, 16, 16, 16, 1, 0.5, None, None, None, 0, None) == -5
    z = torch.randn(8, 16)
    for fn, args in [(ssv_b200.SimclrLoss(True, 0.5), (z, z)), (ssv_b200.MocoLoss(), (z, z, z)),
                     (ssv_b200.BarlowLoss(), (z, z)), (ssv_b200.SimSiamLoss(), (z, z)), (ssv_b200.MSELoss(), (z, z)),
                     (ssv_b200.RelicLoss(), (z, z, z)), (ssv_b200.SwavLoss(), (z, z, z))]:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(*args)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ssv_b200.MemoryBank(10, 4)


def test_missing_library_fails_loudly(cabi, monkeypatch):
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", "/nonexistent/libssv_b200.so")
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        cabi.lib()


def test_dropin_signatures_match_reference():
    """ctor kwargs / defaults / forward parameter names of the reference (utils/losses.py:10,51,122,147,156,206;
    models/moco.py:25,31,38; models/swav.py:46,65,70,77)."""
    import ssv_b200 as S

    def sig(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values() if p.name != "self"]
    E = inspect.Parameter.empty
    assert sig(S.SimclrLoss.__init__) == [("normalize", False), ("temperature", 1.0)]
    assert sig(S.SimclrLoss.forward) == [("zi", E), ("zj", E)]
    assert sig(S.MocoLoss.__init__) == [("normalize", True), ("temperature", 1.0)]
    assert sig(S.MocoLoss.forward) == [("query", E), ("keys", E), ("memory_vectors", E)]
    assert sig(S.BarlowLoss.__init__) == [("normalize", True), ("off_diagonal_weight", 0.005)]
    assert sig(S.BarlowLoss.forward) == [("z_i", E), ("z_j", E)]
    assert sig(S.SimSiamLoss.__init__) == []
    assert sig(S.SimSiamLoss.forward) == [("online_output", E), ("target_output", E)]
    assert sig(S.RelicLoss.__init__) == [("normalize", True), ("temperature", 1.0), ("alpha", 0.5)]
    assert sig(S.RelicLoss.forward) == [("zi", E), ("zj", E), ("z_orig", E)]
    assert sig(S.SwavLoss.__init__) == [("temperature", 0.1), ("sinkhorn_eps", 0.05), ("sinkhorn_iters", 3)]
    assert sig(S.SwavLoss.forward) == [("z_1", E), ("z_2", E), ("prototypes", E), ("bank_features", None)]
    assert sig(S.SwavLoss.compute_codes_sinkhorn) == [("scores", E)]
    assert sig(S.MSELoss.forward) == [("input", E), ("target", E)]
    assert [n for n, _ in sig(S.MemoryBank.__init__)][:2] == ["queue_size", "feature_size"]
    assert sig(S.MemoryBank.add_batch) == [("batch", E)] and sig(S.MemoryBank.get_vectors) == []
    assert [n for n, _ in sig(S.FeatureBank.__init__)][:2] == ["bank_size", "feature_dim"]
    assert sig(S.FeatureBank.add_vectors) == [("fvecs", E)] and sig(S.FeatureBank.return_vectors) == [("device", E)]
    assert sig(S.Prototypes.__init__) == [("hidden_dim", E), ("prototype_size", E)]
    # utils/losses.py:77,80,94,100; models/pirl.py:24,32,36,40,43
    assert sig(S.DinoLoss.__init__) == []
    assert sig(S.DinoLoss.forward) == [("teacher_fvecs", E), ("student_fvecs", E), ("temp_s", E), ("temp_t", E), ("center", E)]
    assert sig(S.PirlLoss.__init__) == [("normalize", True), ("temperature", 1.0), ("loss_weight", 0.5)]
    assert sig(S.PirlLoss.forward) == [("img_features", E), ("patch_features", E), ("memory_pos_features", E),
                                       ("memory_neg_features", E)]
    assert sig(S.PirlMemoryBank.__init__)[:4] == [("data_size", E), ("feature_size", E), ("momentum", 0.5),
                                                  ("num_negatives", 1000)]
    assert sig(S.PirlMemoryBank.initialize_vectors) == [("indices", E), ("vectors", E)]
    assert sig(S.PirlMemoryBank.update_vectors) == [("indices", E), ("new_vectors", E)]
    assert sig(S.PirlMemoryBank.get_positives) == [("indices", E)] and sig(S.PirlMemoryBank.get_negatives) == [("exclude_idx", E)]
    # multi-GPU variants keep the single-GPU ctor kwargs in front
    from ssv_b200 import dist as D
    assert sig(D.DistributedSimclrLoss.__init__)[:2] == [("normalize", False), ("temperature", 1.0)]
    assert sig(D.DistributedBarlowLoss.__init__)[:2] == [("normalize", True), ("off_diagonal_weight", 0.005)]
    assert sig(D.DistributedSwavLoss.__init__)[:3] == [("temperature", 0.1), ("sinkhorn_eps", 0.05), ("sinkhorn_iters", 3)]
    assert sig(D.DistributedMocoLoss.__init__)[:2] == [("normalize", True), ("temperature", 1.0)]
    assert sig(D.DistributedMocoLoss.forward) == [("query", E), ("keys", E), ("memory_vectors", E)]
    assert sig(D.ShardedMemoryBank.add_batch) == [("batch", E)] and sig(D.ShardedMemoryBank.get_vectors) == []
    # config-splat construction as in models/simclr.py:58 etc.
    S.SimclrLoss(**{"normalize": True, "temperature": 0.5})
    S.SwavLoss(**{"temperature": 0.1, "sinkhorn_eps": 0.05, "sinkhorn_iters": 3})


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "self-supervised-vision_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"


def test_relocated_package_binds_from_its_own_header(tmp_path):
    """the package must not depend on the repository layout - `make` ships the public header next to
    the library, and a copy of ssv_b200/ placed anywhere binds every prototype from it."""
    import shutil
    import subprocess
    src = os.path.join(ROOT, "self-supervised-vision_b200", "ssv_b200")
    dst = tmp_path / "site" / "ssv_b200"
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
    assert (dst / "ssv_b200.h").exists(), "make must place the header next to the library"
    code = ("import sys; sys.path.insert(0, sys.argv[1]); import ssv_b200; from ssv_b200 import _cabi;"
            "assert _cabi.HEADER_PATH.startswith(sys.argv[1]), _cabi.HEADER_PATH;"
            "L = _cabi.lib(); assert L.ssvb_version() >= 100; print(len(_cabi.parse_header()))")
    out = subprocess.run([sys.executable, "-c", code, str(tmp_path / "site")], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-1500:]
    assert int(out.stdout.strip()) >= 80
