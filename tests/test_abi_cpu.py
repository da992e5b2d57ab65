"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol
`include/vqb200.h` declares; the Python mirror keeps the reference's constructors, state_dict keys and
error behaviour.  No compute calls (there is no GPU here)."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "vqb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vqb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import vqb200
    lib = vqb200._lib.load()
    declared = _header_symbols()
    assert declared, "header parse found nothing"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vqb200.h but not exported"
    assert sorted(vqb200._lib.exported_symbols()) == declared, "ctypes table and header disagree"
    assert lib.vqb200_abi_version() == 1
    assert isinstance(vqb200._lib.last_error(), str)
    raw = ctypes.CDLL(vqb200._lib.library_path())
    assert raw.vqb200_abi_version() == 1


def test_size_queries_need_no_gpu():
    import vqb200
    lib = vqb200._lib.load()
    # split-bf16 image = 8 code tiles x (16 KiB hi + 16 KiB lo) + 1024 fp32 values of -|E|^2/2
    split = 8 * 32768 + 1024 * 4
    assert lib.vqb200_codebook_image_bytes(1000, 24) == split
    # D == 64 appends (1 KiB aligned) the fp16 filter image: 8 x 16 KiB tiles + 8 x 528 B meta records, and (1 KiB
    # aligned) the group-interleaved fp32 copy for the exact re-rank
    f16 = -(-split // 1024) * 1024 + 8 * 16384 + 8 * 528
    assert lib.vqb200_codebook_image_bytes(1024, 64) == -(-f16 // 1024) * 1024 + 1024 * 64 * 4
    assert lib.vqb200_unique_workspace_bytes() > 256 * 1024
    assert lib.vqb200_assign_workspace_bytes(1000, 64) >= 1000 * 4
    assert lib.vqb200_assign_workspace_bytes(1000, 256) >= 1024 * 256 * 4   # + split-bf16 row image


def test_argument_errors_do_not_need_a_gpu():
    import vqb200
    lib = vqb200._lib.load()
    rc = lib.vqb200_vq_metrics(None, 4, 4, None, 4, 0.25, 1, None, None)
    assert rc == -1 and "null" in vqb200._lib.last_error()
    rc = lib.vqb200_fsq_forward(None, 1, 99, 1, None, 1000, None, None, None, None, None)
    assert rc < 0
    # peer-memory exchange (csrc/peer.cu): argument checks come before any CUDA call
    import ctypes
    dummy = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    assert lib.vqb200_peer_barrier(dummy, 0, 17, ctypes.c_uint32(1), None) == -2          # > VQB200_MAX_PEERS ranks
    assert lib.vqb200_peer_barrier(dummy, 2, 2, ctypes.c_uint32(1), None) == -2           # rank outside the world
    assert lib.vqb200_peer_barrier(None, 0, 1, ctypes.c_uint32(1), None) == -1
    assert lib.vqb200_peer_open(None, None) == -1 and lib.vqb200_peer_alloc(0, None, None) == -1
    assert lib.vqb200_ema_finalize_peer(dummy, dummy, 0, 2, ctypes.c_uint32(1), None, None, None, None, 8, 8,
                                        0.99, 1e-5, None, None, None, None, None) == -1
    assert lib.vqb200_peer_close(None) == 0 and lib.vqb200_peer_free(None) == 0           # NULL is a no-op


def test_single_launch_rvq_size_queries_need_no_gpu():
    import ctypes
    import vqb200
    lib = vqb200._lib.load()
    K = (ctypes.c_int64 * 4)(512, 512, 512, 512)
    assert lib.vqb200_rvq_small_eligible(512, 64, 4, K) == 1
    assert lib.vqb200_rvq_small_eligible(512, 32, 4, K) == 0 and lib.vqb200_rvq_small_eligible(5000, 64, 4, K) == 0
    # per stage [dw (K*64) | cnt (K)] padded to 16 bytes; the workspace adds K + 8 scratch floats per stage + 16
    assert lib.vqb200_rvq_small_stats_floats(4, K) == 4 * 512 * 65
    assert lib.vqb200_rvq_small_workspace_floats(4, K) == 4 * 512 * 65 + 4 * (512 + 8) + 16
    K_odd = (ctypes.c_int64 * 2)(333, 7)
    assert lib.vqb200_rvq_small_stats_floats(2, K_odd) == (333 * 65 + 3) // 4 * 4 + (7 * 65 + 3) // 4 * 4


def test_state_dict_keys_match_reference():
    from models.vqvae import DualMotionVQVAE
    table = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    for key, ref in table.items():
        arch, method = key.split("/")
        if arch == "transformer" and method not in ("hybrid", "ema"):
            continue                                  # same encoder keys; keep the CPU suite quick
        m = DualMotionVQVAE(human_input_dim=126, robot_input_dim=29, hidden_dim=64, arch=arch, method=method,
                            window_size=10)
        mine = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]
        assert mine == ref, key


def test_same_seed_same_initial_state_as_reference_order():
    """Constructors consume the RNG in the reference's order (models/vqvae.py:19-26)."""
    import vqb200
    torch.manual_seed(3)
    q = vqb200.VectorQuantizer(32, 8, use_ema=True)
    torch.manual_seed(3)
    emb = torch.nn.Embedding(32, 8)
    emb.weight.data.uniform_(-1 / 32, 1 / 32)
    ema_w = torch.empty(32, 8).normal_()
    assert torch.equal(q.embedding.weight, emb.weight) and torch.equal(q.ema_w, ema_w)
    assert q.decay == 0.99 and not hasattr(vqb200.VectorQuantizer(4, 4), "decay")


def test_unknown_method_and_cpu_input_fail_loudly():
    import vqb200
    from models.vqvae import DualMotionVQVAE
    with pytest.raises(ValueError):
        DualMotionVQVAE(method="nope", arch="resnet_no_down")
    for mod in (vqb200.VectorQuantizer(8, 4), vqb200.ResidualVQ(2, 8, 4), vqb200.FSQ([8, 5, 5, 5], 4, 4),
                vqb200.LFQ(4), vqb200.HybridVQ(4, vq_codebook_size=8)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            mod(torch.randn(2, 4, 3))
    loss, z, met = vqb200.IdentityVQ()(torch.randn(2, 4, 3))
    assert float(loss) == 0.0 and float(met["perplexity"]) == 1.0


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package or models/ may import it."""
    pkg = os.path.join(ROOT, "bridging-the-gap-of-robot-learning-via-distribution-reinforcement-learning-vq-vae_b200")
    for base in (pkg, os.path.join(ROOT, "models")):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_token_layout_and_file_format_need_no_gpu(tmp_path):
    """Token spec per module, the C-ABI size query, the oracle's bit loop and the file round trip (all on CPU)."""
    import numpy as np
    import vqb200
    from vqb200 import tokens
    from oracle.tokens_oracle import pack, unpack
    lib = vqb200._lib.load()
    hy = tokens.spec_for(vqb200.HybridVQ(64, [8, 5, 5, 5], 512), 7, 64, 3)
    assert (hy.S, hy.code_bits, hy.d, hy.digit_bits, hy.bytes_per_token) == (4, 9, 4, 8, 9)
    assert tokens.spec_for(vqb200.ResidualVQ(4, 1024, 64), 7, 64, 3).bytes_per_token == 5
    assert tokens.spec_for(vqb200.VectorQuantizer(1000, 64), 7, 64, 3).code_bits == 10
    assert tokens.spec_for(vqb200.FSQ([8, 5, 5, 5], 64, 64), 7, 64, 3, digit_bits=6).bytes_per_token == 3
    assert tokens.spec_for(vqb200.LFQ(64, 10), 7, 64, 3).bytes_per_token == 2
    with pytest.raises(RuntimeError):
        tokens.spec_for(vqb200.IdentityVQ(), 1, 64, 1)
    assert lib.vqb200_token_bytes(4, 9, 4, 8) == 9
    assert lib.vqb200_token_bytes(0, 0, 0, 0) < 0 and lib.vqb200_token_bytes(8, 32, 16, 32) < 0     # empty / > 256 bits
    # oracle bit loop is its own inverse, including negative digits
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 512, (4, 21))
    digits = rng.integers(-128, 128, (21, 4))
    t = pack(codes, digits, 9, 8)
    c2, d2 = unpack(t, 4, 9, 4, 8)
    assert np.array_equal(c2, codes) and np.array_equal(d2, digits)
    # file round trip
    tb = tokens.TokenBatch(hy, torch.from_numpy(rng.integers(0, 256, (21, 9), dtype=np.uint8)))
    f = str(tmp_path / "x.vqtok")
    tokens.save(f, tb)
    back = tokens.load(f)
    assert back.spec == hy and torch.equal(back.data, tb.data)
    with pytest.raises(RuntimeError):
        tokens.decode(vqb200.HybridVQ(64, [8, 5, 5, 5], 512), back)      # CPU tokens: no CPU path
    # a token header is untrusted input: a layout that disagrees with its own fields, or with the module, is rejected
    # before any kernel could index with it
    import dataclasses, struct
    bad = dataclasses.replace(hy, bytes_per_token=hy.bytes_per_token - 1)
    raw = json.dumps(dataclasses.asdict(bad)).encode()
    g = str(tmp_path / "bad.vqtok")
    with open(g, "wb") as fh:
        fh.write(tokens.MAGIC + struct.pack("<I", len(raw)) + raw + bytes(21 * (hy.bytes_per_token - 1)))
    with pytest.raises(RuntimeError, match="bytes per token"):
        tokens.load(g)
    with pytest.raises(RuntimeError, match="embedding_dim"):
        tokens.decode(vqb200.HybridVQ(32, [8, 5, 5, 5], 512), back)      # tokens are for C = 64
    with pytest.raises(RuntimeError, match="FSQ digits"):
        tokens.decode(vqb200.HybridVQ(64, [8, 5, 5], 512), back)         # 3 digits instead of 4


def test_library_staleness_is_detected_by_content(tmp_path, monkeypatch):
    """_lib.load() rebuilds when the sources changed since the library was linked; the check hashes file contents
    (mtimes do not survive a copy of the tree to the GPU box)."""
    from vqb200 import build as b
    if not os.path.exists(b.LIB):
        pytest.skip("library not built")
    assert os.path.exists(b.STAMP) and not b.needs_build()
    monkeypatch.setenv("VQB200_NVCC_EXTRA", "-DVQB200_SOME_VARIANT=1")    # flags are part of the hash
    assert b.needs_build()


def test_staged_reference_is_unmodified_when_present():
    """oracle/_ref (git-ignored, staged by build() from /root/reference) must hold byte-identical copies."""
    import hashlib, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "MANIFEST.json")):
        import pytest
        pytest.skip("oracle/_ref not staged in this checkout")
    m = json.load(open(os.path.join(ref, "MANIFEST.json")))
    assert "models/vqvae.py" in m["files"] and "scripts/train_ablation.py" in m["files"]
    for rel, digest in m["files"].items():
        assert hashlib.sha256(open(os.path.join(ref, rel), "rb").read()).hexdigest() == digest, rel
        src = os.path.join(m["source"], rel)
        if os.path.exists(src):          # only in the build container
            assert hashlib.sha256(open(src, "rb").read()).hexdigest() == digest, f"{rel} differs from the reference tree"


def test_reference_arm_runs_the_staged_reference_modules():
    """bench.py's CPU arm: the unmodified reference ResidualVQ (torch CPU) on a tiny sample; kind == 'reference'."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "models", "vqvae.py")):
        import pytest
        pytest.skip("oracle/_ref not staged in this checkout")
    sys.path.insert(0, root)
    import bench
    r = bench.cpu_reference_arm(dict(kind="rvq", S=2, K=64, D=64, B=64, T=10), steps=1, warmup=1, chunk_vectors=640)
    assert r["kind"] == "reference" and r["value"] > 0 and r["cores"] >= 1
