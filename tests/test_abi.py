"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/szn.h declares;
the module surface matches the reference's (no compute calls: there is no GPU here)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "szn.h")).read()
    return sorted(set(re.findall(r"\b(szn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from zeroshotsemanticsegmentation_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libszn.so does not export %s" % n
    # every bound signature is declared in the header and vice versa
    assert set(_lib.SIGNATURES) | {"szn_last_error", "szn_launch_count", "szn_abi_version",
                                   "szn_embed_argmax_scratch_floats", "szn_head_fused_workspace_floats",
                                   "szn_comm_available"} == set(names)


def test_library_is_built_from_the_sources_in_the_tree():
    """include/szn_build.h: the library reports the hash of the sources it was compiled from; the Makefile computes the same
    hash from the files in the tree.  nvcc objects are not byte-reproducible, so this -- not a hash of the .so -- is what
    ties a profile (profiles/r02_umma_traffic.json: build_id) to a build."""
    import subprocess
    import __graft_entry__ as ge
    ge.build()
    from zeroshotsemanticsegmentation_b200 import _lib
    want = subprocess.check_output(["make", "-s", "-C", os.path.join(ROOT, "zeroshotsemanticsegmentation_b200", "csrc"),
                                    "print-hash"], text=True).strip()
    assert len(want) == 16 and _lib.build_id() == want
    hdr = open(os.path.join(ROOT, "include", "szn_build.h")).read()
    assert "szn_build_id" in hdr


def test_no_cpu_fallback():
    import zeroshotsemanticsegmentation_b200 as szn
    m = szn.FCN32s(4)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        szn.utils.cosine_loss(torch.zeros(1, 4, 2, 2), torch.zeros(1, 2, 2, dtype=torch.long),
                              torch.zeros(1, 4, 2, 2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "zeroshotsemanticsegmentation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no oracle", "")


def test_module_surface_matches_reference():
    import torch.nn as nn
    import zeroshotsemanticsegmentation_b200 as szn
    from oracle import szn_oracle as O
    m = szn.FCN32s(20)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in O.init_params(20).items()}
    assert torch.equal(m.upscore.weight.detach(), O.upsampling_weight(20, 20))
    vgg = nn.Module()
    feats = []
    for row in O.TRUNK:
        feats.append(nn.MaxPool2d(2) if len(row) == 1 else nn.Conv2d(row[1], row[2], 3, padding=1))
    vgg.features = nn.Sequential(*feats)
    vgg.classifier = nn.Sequential(nn.Linear(25088, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 4096))
    m.copy_params_from_vgg16(vgg)
    assert torch.equal(m.conv3_2.weight, vgg.features[7].weight)
    assert torch.equal(m.fc6.weight.reshape(4096, -1), vgg.classifier[0].weight)
    assert m.fc6.weight.is_contiguous(memory_format=torch.channels_last)  # back in the layout the wgrad kernel writes


def test_host_metrics_equal_live_reference():
    """utils.label_accuracy_score (numpy path) against the unmodified reference, when its checkout is present."""
    import numpy as np
    from oracle import ref_import
    if ref_import.reference_root() is None:
        pytest.skip("reference checkout not present (GPU box)")
    _, RU = ref_import.load_reference()
    import zeroshotsemanticsegmentation_b200 as szn
    rng = np.random.RandomState(3)
    for n_class, unseen in ((21, [3, 17]), (33, None)):
        lt = [rng.randint(-1, n_class, (40, 30)) for _ in range(2)]
        lp = [np.where(rng.rand(40, 30) < 0.7, np.clip(t, 0, None), rng.randint(0, n_class, (40, 30))) for t in lt]
        a = szn.utils.label_accuracy_score(lt, lp, n_class, unseen)
        b = RU.label_accuracy_score(lt, lp, n_class, unseen)
        np.testing.assert_allclose(np.array(a, dtype=np.float64), np.array(b, dtype=np.float64), rtol=1e-12)


def test_checkpoint_dict_round_trip(tmp_path):
    """The reference's checkpoint format (trainer_fcn.py:281-292) and resume path (train.py:110-116,135-136):
    a dict with model_state_dict / optim_state_dict, loaded with load_state_dict(strict=False)."""
    import zeroshotsemanticsegmentation_b200 as szn
    from oracle import szn_oracle as O
    m = szn.FCN32s(n_class=20)
    m.load_state_dict(O.init_params(20, seed=3))
    optim = torch.optim.SGD([{"params": [m.conv1_2.weight]}, {"params": [m.conv1_2.bias], "lr": 2e-3, "weight_decay": 0}],
                            lr=1e-3, momentum=0.99, weight_decay=5e-4)
    path = str(tmp_path / "checkpoint")
    torch.save({"epoch": 3, "iteration": 1234, "arch": m.__class__.__name__, "optim_state_dict": optim.state_dict(),
                "model_state_dict": m.state_dict(), "best_mean_iu": 0.25}, path)
    ck = torch.load(path)
    assert ck["arch"] == "FCN32s"
    m2 = szn.FCN32s(n_class=20)
    missing = m2.load_state_dict(ck["model_state_dict"], strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k
    # the kernel-friendly (channels_last) parameter layout survives loading; shapes and values are the reference's
    assert m2.conv3_2.weight.is_contiguous(memory_format=torch.channels_last) and m2.conv3_2.weight.shape == (256, 256, 3, 3)
    # a reference checkpoint has exactly these 36 tensors
    assert len(ck["model_state_dict"]) == 36


def test_seen_unseen_helpers_match_oracle():
    """utils.split_embeddings / seenmask_target vs the oracle's restatement of trainer_fcn.py:44-64 and
    trainer_seenmask.py:55-56 (host logic, any device)."""
    import zeroshotsemanticsegmentation_b200 as szn
    from oracle import szn_oracle as O
    g = torch.Generator().manual_seed(5)
    table = torch.randn(21, 20, generator=g)
    unseen = [3, 17, 20]
    a, b = szn.utils.split_embeddings(table, unseen)
    ra, rb = O.split_tables(table, unseen)
    assert torch.equal(a, ra) and torch.equal(b, rb)
    t = torch.randint(-1, 21, (2, 9, 7), generator=g)
    assert torch.equal(szn.utils.seenmask_target(t, unseen, 21), O.seenmask_target(t, unseen, 21))


def test_missing_extension_fails_loudly():
    """No silent fallback: without libszn.so the binding raises (checked in a fresh interpreter)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from zeroshotsemanticsegmentation_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept RuntimeError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n" % ROOT)
    env = dict(os.environ, SZN_LIB="/nonexistent/libszn.so")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "RAISED True" in out.stdout, out.stdout + out.stderr


def _c_prototypes():
    """name -> list of C parameter type strings, parsed from include/szn.h."""
    src = open(os.path.join(ROOT, "include", "szn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for ret, name, params in re.findall(r"\b(int|long long|const char\*)\s+(szn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        params = params.strip()
        types = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                types.append(p.rsplit(" ", 1)[0] if not p.endswith("*") else p)  # drop the parameter name
        protos[name] = (ret, types)
    return protos


def test_ctypes_signatures_match_the_header_prototypes():
    """Every argtypes list of the ctypes binding has the arity and the scalar widths of the C prototype it binds
    (a mismatch would corrupt arguments silently: ctypes cannot check it at run time)."""
    import ctypes
    from zeroshotsemanticsegmentation_b200 import _lib
    protos = _c_prototypes()

    def expected(ctype):
        if "*" in ctype:
            return ctypes.c_void_p
        return {"int": ctypes.c_int, "long long": ctypes.c_longlong, "unsigned long long": ctypes.c_ulonglong,
                "float": ctypes.c_float}[ctype]

    assert set(_lib.SIGNATURES) <= set(protos)
    for name, argtypes in _lib.SIGNATURES.items():
        ret, ctypes_c = protos[name]
        assert ret == "int", name
        want = [expected(t) for t in ctypes_c]
        assert list(argtypes) == want, "%s: binding %s vs header %s" % (name, argtypes, ctypes_c)
