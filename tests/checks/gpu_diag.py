"""Staged GPU diagnostics (development aid; the judged checks are tests/ -m gpu).
    python tests/checks/gpu_diag.py <stage> ...      stages: gemm query map enc_tiny enc_full text timing
Each stage prints max errors against the CPU oracle; run each under `timeout` on the GPU box."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ovo_b200 import synth  # noqa: E402
from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict, gemm_bf16  # noqa: E402
from ovo_b200.map import SemanticMap  # noqa: E402
from oracle import encoder as OE, fusion as OF  # noqa: E402

dev = "cuda"


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), (a - b).abs().max().item()


def stage_gemm():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 128, 64), (128, 256, 128), (300, 200, 192), (1154, 3072, 1024), (1154, 1024, 4096),
                      (1152, 1024, 640), (77, 1024, 1024), (100000, 20, 1024)]:
        A = torch.randn(M, K, device=dev).bfloat16()
        B = torch.randn(N, K, device=dev).bfloat16()
        bias = torch.randn(N, device=dev)
        ref = A.float() @ B.float().T + bias
        for bn in ([32, 64, 128, 256] if N > 32 else [32]):
            if N <= 128 and bn > 128 and N <= 32:
                continue
            out = gemm_bf16(A, B, bias, force_bn=bn)
            torch.cuda.synchronize()
            r, m = relerr(out, ref)
            print(f"gemm M={M} N={N} K={K} bn={bn}: rel {r:.3e} max {m:.3e}", flush=True)
        out = gemm_bf16(A, B, None)
        r, m = relerr(out, ref - bias)
        print(f"gemm M={M} N={N} K={K} auto: rel {r:.3e} max {m:.3e}", flush=True)


def stage_query():
    torch.manual_seed(0)
    sm = SemanticMap()
    for (N, Q) in [(5000, 20), (100000, 21), (50000, 200), (20000, 300)]:
        bank = torch.nn.functional.normalize(torch.randn(N, 1024, device=dev), dim=-1).bfloat16()
        text = torch.nn.functional.normalize(torch.randn(Q, 1024, device=dev), dim=-1)
        out = sm.query_dense(bank, text)
        ref = bank.float() @ text.bfloat16().float().T
        r, m = relerr(out, ref)
        ref32 = bank.float() @ text.T
        r2, m2 = relerr(out, ref32)
        print(f"query_dense N={N} Q={Q}: vs bf16-text rel {r:.3e} max {m:.3e}; vs f32-text max {m2:.3e}", flush=True)
        cls, conf = sm.classify(out, 0.0)
        rc = out.argmax(1)
        print("  classify mismatches", int((cls.long() != torch.where(out.max(1).values > 0, rc, -1)).sum()))
    bank = torch.randn(37, 1024, device=dev)
    text = torch.randn(5, 1024, device=dev)
    r, m = relerr(sm.query_instances(bank, text), bank @ text.T)
    print(f"query_instances rel {r:.3e} max {m:.3e}")


def stage_map():
    sm = SemanticMap()
    K = synth.intrinsics()
    next_id = 0
    N = 200000
    d0 = synth.depth_map(frame_id=0)
    xyz, ids, ins = synth.point_map(N, d0, K, synth.pose(0), seed=0, frac_visible=0.5)
    ins_o = ins.copy()
    xyz_d = torch.from_numpy(xyz).to(dev)
    ins_d = torch.from_numpy(ins).to(dev)
    seg, bm = synth.grid_masks()
    seg_d = torch.from_numpy(seg).to(dev)
    bank = torch.zeros(N, 256, device=dev, dtype=torch.bfloat16)
    bank_lo = torch.zeros_like(bank)
    counts = torch.zeros(N, device=dev, dtype=torch.int32)
    bank_o = np.zeros((N, 256), np.float32); lo_o = np.zeros_like(bank_o); counts_o = np.zeros(N, np.int32)
    for fid in range(4):
        c2w = synth.pose(fid * 5)
        d = synth.depth_map(frame_id=fid)
        dd = torch.from_numpy(d).to(dev)
        df = sm.depth_filter(dd).cpu().numpy()
        print("depth_filter mismatches", int((df != OF.depth_filter(d)).sum()))
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        t0 = time.time()
        votes, nm, nxt = sm.associate(xyz_d, ins_d, dd, seg_d, c2w, K, next_id, kf_slot=fid)
        t1 = time.time()
        seg_of_pt, fm = OF.associate(xyz, ins_o, d, seg, c2w, w2c, K, 0.05, True)
        ins_new, rows, nxt_o = OF.track(ins_o, seg_of_pt, seg, 100, next_id)
        bad = {k: int((votes[k] != np.array([r[k] for r in rows])).sum()) for k in votes}
        print(f"frame {fid}: n_matched {nm} vs {(seg_of_pt > -2).sum()}  next {nxt} vs {nxt_o}  row mismatches {bad}  "
              f"ins mismatches {int((ins_d.cpu().numpy() != ins_new).sum())}  t={1e3 * (t1 - t0):.2f} ms", flush=True)
        # dense fusion
        order, fused, mask_row = OF.fuse_masks(bm, rows)
        R = max(len(order), 1)
        feats = torch.randn(R, 256)
        sm.fuse_dense(fid, bank, bank_lo, counts, feats.to(dev), torch.from_numpy(mask_row).to(dev))
        OF.dense_fuse(bank_o, lo_o, counts_o, [np.where(seg_of_pt >= 0, seg_of_pt, -1)], [mask_row], feats.numpy())
        print("  dense bank mismatches", int((bank.float().cpu().numpy() != bank_o).sum()) + int((bank_lo.float().cpu().numpy() != lo_o).sum()),
              "count mismatches", int((counts.cpu().numpy() != counts_o).sum()), flush=True)
        ins_o, next_id = ins_new, nxt_o


def _enc(cfg, n_img=2, seed=0, text=True, **kw):
    sd = random_state_dict(cfg, seed=seed, text=text)
    enc = RegionEncoder(cfg, sd, max_images=max(n_img, 2), **kw)
    ocfg = OE.VitCfg(**{k: getattr(cfg, k) for k in ("image_size", "patch_size", "width", "layers", "heads", "mlp_width",
                                                      "output_dim", "ln_eps", "text_ctx", "text_width", "text_heads",
                                                      "text_layers", "text_mlp_width", "vocab_size")})
    return enc, sd, ocfg


def _enc_check(cfg, layers_list, do_regions=True):
    enc, sd, ocfg = _enc(cfg)
    torch.manual_seed(1)
    px = torch.randn(2, 3, cfg.image_size, cfg.image_size) * 0.5
    for L in layers_list:
        out = enc.forward_features_from_pixels(px.to(dev), n_layers=L, ln_post=(L == cfg.layers))
        torch.cuda.synchronize()
        with torch.no_grad():
            ref = OE.vit_forward_features(px, sd, ocfg, n_layers=L, norm=(L == cfg.layers))
        r, m = relerr(out, ref)
        cos = torch.nn.functional.cosine_similarity(out.cpu().flatten(0, 1), ref.flatten(0, 1), dim=-1).min().item()
        print(f"layers={L}: rel-L2 {r:.3e} max {m:.3e} min token cos {cos:.6f} nan={bool(torch.isnan(out).any())}", flush=True)
    if do_regions:
        img = synth.rgb(480, 640, seed=3)
        seg, bm = synth.grid_masks(480, 640, rows=3, cols=4)
        bm = np.concatenate([bm, np.zeros((1, 480, 640), bool)])
        bm[-1, 200:203, 5:9] = True      # tiny mask -> no token -> NaN row in the reference
        out = enc.encode_regions(torch.from_numpy(img).to(dev), torch.from_numpy(bm).to(dev))
        torch.cuda.synchronize()
        with torch.no_grad():
            ref = OE.encode_regions(img, bm, sd, ocfg)
        ok = ~torch.isnan(ref).any(1)
        cos = torch.nn.functional.cosine_similarity(out.cpu()[ok], ref[ok], dim=-1)
        r, m = relerr(out.cpu()[ok], ref[ok])
        print(f"regions: min cos {cos.min().item():.6f} rel-L2 {r:.3e} max {m:.3e}; NaN rows ours {torch.isnan(out).any(1).cpu().tolist().count(True)} ref {(~ok).sum().item()}", flush=True)
    return enc, sd, ocfg


def stage_enc_tiny():
    cfg = EncoderConfig(width=128, layers=2, heads=2, mlp_width=512, output_dim=64, text_width=128, text_heads=2,
                        text_layers=2, text_mlp_width=512, text_output_dim=64, vocab_size=1000)
    enc, sd, ocfg = _enc_check(cfg, [0, 1, 2])
    tok = torch.randint(1, 999, (5, 32)); tok[:, 10:] = 0; tok[torch.arange(5), torch.tensor([3, 5, 7, 9, 9])] = 999
    out = enc.encode_text(tok.to(dev))
    with torch.no_grad():
        ref = OE.text_forward(tok, sd, ocfg)
    r, m = relerr(out, ref)
    print(f"text: rel-L2 {r:.3e} max {m:.3e}", flush=True)


def stage_enc_full():
    cfg = EncoderConfig(text_layers=0)
    _enc_check(cfg, [0, 1, 2, 24])


def stage_text():
    cfg = EncoderConfig(layers=1)
    enc, sd, ocfg = _enc(cfg)
    tok = torch.randint(1, 49000, (21, 32)); tok[:, 12:] = 0; tok[:, 11] = 49407
    out = enc.encode_text(tok.to(dev))
    with torch.no_grad():
        ref = OE.text_forward(tok, sd, ocfg)
    r, m = relerr(out, ref)
    cos = torch.nn.functional.cosine_similarity(out.cpu(), ref, dim=-1).min().item()
    print(f"text L14: rel-L2 {r:.3e} max {m:.3e} min cos {cos:.6f}", flush=True)


def _time(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def stage_timing():
    sm = SemanticMap()
    for (N, Q) in [(2_000_000, 20), (2_000_000, 200)]:
        bank = torch.randn(N, 1024, device=dev).bfloat16()
        text = torch.randn(Q, 1024, device=dev)
        out = torch.empty(N, Q, device=dev)
        ms = _time(lambda: sm.query_dense(bank, text, out), n=10)
        gb = (N * 1024 * 2 + N * Q * 4 + Q * 1024 * 2) / 1e9
        print(f"query_dense N={N} Q={Q}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s", flush=True)
        del bank, out
    for M, N, K in [(1154, 3072, 1024), (1154, 4096, 1024), (1154, 1024, 4096), (9232, 3072, 1024), (9232, 4096, 1024), (9232, 1024, 4096)]:
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
        for bn in (128, 256):
            ms = _time(lambda: gemm_bf16(A, B, None, force_bn=bn), n=20)
            print(f"gemm {M}x{N}x{K} bn={bn}: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
        ms = _time(lambda: A.float() @ B.float().T if False else torch.matmul(A, B.T), n=20)
        print(f"  cublas: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    cfg = EncoderConfig(text_layers=0)
    for n_img in (2, 8, 16):
        enc, sd, _ = _enc(cfg, n_img=n_img, text=False)
        px = torch.randn(n_img, 3, 336, 336, device=dev)
        ms = _time(lambda: enc.forward_features_from_pixels(px), n=10)
        print(f"vit forward n_img={n_img}: {ms:.3f} ms  {n_img * 349.2 / ms:.1f} TFLOP/s", flush=True)
        del enc
    K = synth.intrinsics(); d = synth.depth_map(); N = 2_000_000
    xyz, ids, ins = synth.point_map(N, d, K, synth.pose(0), seed=0)
    seg, bm = synth.grid_masks()
    xyz_d, ins_d, dd, seg_d = (torch.from_numpy(a).to(dev) for a in (xyz, ins, d, seg))
    ms = _time(lambda: sm.associate(xyz_d, ins_d.clone(), dd, seg_d, synth.pose(0), K, 0), n=10)
    print(f"associate N={N}: {ms:.3f} ms", flush=True)


def stage_clusters():
    """GEMM correctness + speed per thread-block-cluster size (TMA multicast of the weight tile)."""
    from ovo_b200 import _lib
    torch.manual_seed(0)
    for cs in (1, 2, 4):
        _lib.lib().ovo_set_gemm_cluster(cs)
        for (M, N, K) in [(300, 512, 192), (1154, 3072, 1024), (9232, 1024, 4096), (9232, 4096, 1024), (9232, 3072, 1024), (9232, 1024, 1024)]:
            A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
            bias = torch.randn(N, device=dev)
            ref = A.float() @ B.float().T + bias
            for bn in (128, 256):
                out = gemm_bf16(A, B, bias, force_bn=bn)
                torch.cuda.synchronize()
                r, m = relerr(out, ref)
                ms = _time(lambda: gemm_bf16(A, B, bias, force_bn=bn), n=20)
                print(f"cs={cs} gemm {M}x{N}x{K} bn={bn}: rel {r:.2e}  {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    cfg = EncoderConfig(text_layers=0)
    enc, sd, ocfg = _enc(cfg, n_img=16, text=False)
    px = torch.randn(16, 3, 336, 336, device=dev)
    for cs in (1, 2, 4, 0):
        _lib.lib().ovo_set_gemm_cluster(cs)
        os.environ["X"] = "1"
        enc.lib.ovo_profile_begin()
        for _ in range(2):
            enc.forward_features_from_pixels(px)
        torch.cuda.synchronize()
        prof = _lib.profile_report()
        print(f"cs={cs} vit16 profiled: " + ", ".join(f"{k} {v['ms'] / 2:.3f} ms" for k, v in prof.items() if v['launches']), flush=True)
    _lib.lib().ovo_set_gemm_cluster(0)
    out = enc.forward_features_from_pixels(px[:2])
    with torch.no_grad():
        ref = OE.vit_forward_features(px[:2].cpu(), sd, ocfg)
    r, m = relerr(out, ref)
    print(f"vit parity after cluster runs: rel-L2 {r:.3e}", flush=True)
    for n_img in (2, 16):
        ms = _time(lambda: enc.forward_features_from_pixels(px[:n_img]), n=10)
        print(f"vit forward (graph) n_img={n_img}: {ms:.3f} ms  {n_img * 349.2 / ms:.1f} TFLOP/s", flush=True)


def stage_querycs():
    """Dense query at Q=200 with the text tile multicast over clusters of 1/2/4 CTAs."""
    from ovo_b200 import _lib
    sm = SemanticMap()
    N = 2_000_000
    bank = torch.randn(N, 1024, device=dev).bfloat16()
    for Q in (200, 64, 128):
        text = torch.randn(Q, 1024, device=dev)
        out = torch.empty(N, Q, device=dev)
        ref = None
        for cs in (1, 2, 4):
            _lib.lib().ovo_set_gemm_cluster(cs << 16)
            sm.query_dense(bank, text, out)
            torch.cuda.synchronize()
            if ref is None:
                ref = out[:4096].clone()
            same = torch.equal(ref, out[:4096])
            ms = _time(lambda: sm.query_dense(bank, text, out), n=10)
            gb = (N * 1024 * 2 + N * Q * 4 + Q * 1024 * 2) / 1e9
            print(f"query Q={Q} cluster={cs}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  same={same}", flush=True)
    _lib.lib().ovo_set_gemm_cluster(1 << 16)


def stage_attndbg():
    """Where the attention kernel's time goes (attention.cuh dbg bits); results are wrong in the dbg modes."""
    from ovo_b200 import _lib
    cfg = EncoderConfig(text_layers=0, layers=4)
    enc, sd, ocfg = _enc(cfg, n_img=16, text=False)
    px = torch.randn(16, 3, 336, 336, device=dev)
    for dbg, name in ((0, "full"), (1, "no softmax math / P stores"), (2, "no MMA"), (3, "neither (barrier skeleton)"), (4, "one key block"), (7, "one block, skeleton")):
        _lib.lib().ovo_set_gemm_cluster((dbg << 24) | 1)
        _lib.profile_begin()
        for _ in range(2):
            enc.forward_features_from_pixels(px)
        torch.cuda.synchronize()
        prof = _lib.profile_report()
        print(f"{name:32s} 4 layers x16 img: attention {prof['attention']['ms'] / 2:.3f} ms", flush=True)
    _lib.lib().ovo_set_gemm_cluster(1)


def stage_gemmdbg():
    """Which part bounds the GEMM: full vs no-epilogue-stores vs no-MMA vs no-TMA (EpiParams::debug bits)."""
    from ovo_b200 import _lib
    cfg = EncoderConfig(text_layers=0, layers=4)
    enc, sd, ocfg = _enc(cfg, n_img=16, text=False)
    px = torch.randn(16, 3, 336, 336, device=dev)
    for dbg, name in ((0, "full"), (1, "no-epilogue"), (8, "no-gelu-math"), (16, "no-bf16-stores"), (32, "no-bias"), (56, "no gelu/stores/bias"), (2, "no-mma"), (4, "no-tma")):
        _lib.lib().ovo_set_gemm_cluster((dbg << 8) | 1)
        _lib.profile_begin()
        for _ in range(2):
            enc.forward_features_from_pixels(px)
        torch.cuda.synchronize()
        prof = _lib.profile_report()
        print(f"{name:22s} 4 layers x16 img: gemm {prof['gemm']['ms'] / 2:.3f} ms ({prof['gemm']['launches'] // 2} launches)", flush=True)
        for (M, N, K) in [(9232, 4096, 1024), (9232, 1024, 4096), (9232, 1024, 1024)]:
            A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
            ms = _time(lambda: gemm_bf16(A, B, None, force_bn=256), n=10)
            print(f"    f32-out gemm {M}x{N}x{K}: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    _lib.lib().ovo_set_gemm_cluster(0)


def stage_gemmshapes():
    """Per-shape GEMM throughput with the fused epilogues (ovo_gemm_bench): ViT-L/14 x16 images and the SAM-2 shapes."""
    import ctypes as C
    from ovo_b200 import _lib
    lib = _lib.lib()
    names = {0: "f32", 1: "bf16", 2: "gelu", 3: "resid", 4: "qkv", 6: "relu"}
    shapes = [("vit out-proj", 3, 9232, 1024, 1024), ("vit fc2", 3, 9232, 1024, 4096), ("vit fc1", 2, 9232, 4096, 1024),
              ("vit qkv~bf16", 1, 9232, 3072, 1024), ("vit qkv", 4, 9232, 3072, 1024),
              ("sam s3 qkv", 1, 4096, 1728, 576), ("sam s3 proj", 3, 4096, 576, 576), ("sam s3 fc1", 2, 4096, 2304, 576),
              ("sam s3 fc2", 3, 4096, 576, 2304), ("sam s1 qkv", 1, 65536, 432, 144), ("sam s1 fc1", 2, 65536, 576, 144),
              ("sam s2 fc1", 2, 16384, 1152, 288), ("sam dec kproj", 1, 1048576, 128, 256), ("sam dec oproj", 3, 1048576, 256, 128),
              ("sam dec up0", 3, 1048576, 256, 256), ("sam dec up1", 2, 4194304, 128, 64)]
    dbgs = [(0, "")]
    if os.environ.get("OVO_DIAG_NOPREFETCH"):
        dbgs.append((64, " no-prefetch"))
    for name, epi, M, N, K in shapes:
        for dbg, tag in dbgs:
            if dbg and epi != 3:
                continue
            lib.ovo_set_gemm_cluster((dbg << 8) | 1)
            row = []
            for bn in (0, 64, 128, 256):
                if bn and bn > max(32, N):
                    continue
                ms = C.c_float(0)
                _lib.check(lib.ovo_gemm_bench(epi, M, N, K, bn, 20, C.byref(ms), _lib.stream_ptr()), "gemm_bench")
                row.append(f"bn{bn or 'auto'} {ms.value * 1e3:7.1f} us {2.0 * M * N * K / ms.value / 1e9:6.0f} TF/s")
            print(f"{name:14s} {names[epi]:5s}{tag} {M}x{N}x{K}: " + " | ".join(row), flush=True)
    lib.ovo_set_gemm_cluster(0)


if __name__ == "__main__":
    for st in sys.argv[1:]:
        print(f"===== {st}", flush=True)
        t = time.time()
        globals()["stage_" + st]()
        torch.cuda.synchronize()
        print(f"===== {st} done in {time.time() - t:.1f}s", flush=True)
