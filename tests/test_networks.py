"""G / D networks on the sgb200 ops against the golden vectors of the reference networks (net_tiny.npz)
and against the CPU oracle.  CPU part: parameter naming; GPU part: forward + the four training phases."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, assert_close, check_phase_grads as _check, phase_tolerances
from oracle import ref_networks as RN

DEV = 'cuda'


@pytest.fixture(scope='module')
def gold():
    z = np.load(os.path.join(GOLDEN, 'net_tiny.npz'))
    meta = json.loads(str(z['meta']))
    return z, meta


def _build(meta, device, **over):
    from sgb200 import training
    c = meta['cfg']
    cfg = training.TrainConfig(img_resolution=c['img_resolution'], z_dim=c['z_dim'], w_dim=c['w_dim'],
                               channel_base=c['channel_base'], channel_max=c['channel_max'], map_layers=c['map_layers'],
                               d_arch=c['d_arch'], mbstd_group_size=c['mbstd_group_size'], batch_gpu=meta['n'],
                               style_mixing_prob=0.0, noise_mode='const', use_ema=False, r1_gamma=meta['r1_gamma'],
                               pl_weight=meta['pl_weight'], pl_decay=meta['pl_decay'], **over)
    G, D = training.build_networks(cfg, device)
    return cfg, G, D


def _load(z, G, D):
    gsd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('G.')}
    dsd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('D.')}
    G.load_state_dict(gsd, strict=True)
    D.load_state_dict(dsd, strict=True)


def test_state_dict_names_match_reference(gold):
    z, meta = gold
    cfg, G, D = _build(meta, 'cpu')
    gref = {k[2:]: tuple(z[k].shape) for k in z.files if k.startswith('G.')}
    dref = {k[2:]: tuple(z[k].shape) for k in z.files if k.startswith('D.')}
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == gref
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == dref
    _load(z, G, D)


@pytest.mark.gpu
@pytest.mark.parametrize('channels_last', [False, True])
def test_forward_matches_reference_golden(gold, channels_last):
    z, meta = gold
    cfg, G, D = _build(meta, DEV, channels_last=channels_last)
    _load(z, G, D)
    G.train(); D.train()
    zz = torch.from_numpy(z['z']).to(DEV)
    with torch.no_grad():
        ws = G.mapping(zz, None, skip_w_avg_update=True)
        img = G.synthesis(ws, noise_mode='const')
        logits = D(img, None)
        G.eval()
        img_fused = G.synthesis(ws, noise_mode='const')       # eval => fused_modconv (grouped conv) branch
    assert_close(ws, torch.from_numpy(z['ws']), 1e-4, 'ws')
    assert_close(img, torch.from_numpy(z['img']), 1e-4, 'img')
    assert_close(img_fused, torch.from_numpy(z['img']), 1e-4, 'img (fused_modconv)')
    assert_close(logits, torch.from_numpy(z['logits']), 1e-4, 'logits')


# arithmetic modes of the training-phase parity tests (same table as tests/test_ref_callers_gpu.py):
#   strict = fp32 FFMA kernels (reference default allow_tf32=False);  tf32 = tcgen05 kind::tf32 forward / dgrad / wgrad, the
#   arithmetic bench.py measures;  fp16 = num_fp16_res=2 + conv_clamp=256 (+ TF32 in the fp32 blocks), against the fp32 golden
# tolerances: helpers.phase_tolerances (strict 2e-4 / 5e-4 on every tensor; tensor-core modes: median <= 1e-2, worst tensor
# <= 2 x what the reference's own cuDNN path shows on the same gradients)
MODES = {
    'strict': dict(tf32=False, over={}),
    'tf32': dict(tf32=True, over={}),
    'fp16': dict(tf32=True, over=dict(num_fp16_res=2, conv_clamp=256.0)),
}


@pytest.mark.gpu
@pytest.mark.parametrize('mode,channels_last', [('strict', False), ('strict', True), ('tf32', True), ('tf32', False), ('fp16', True)])
def test_training_phases_match_reference_golden(gold, mode, channels_last):
    """Gradients of every parameter for Gmain, Dmain, Dreg (R1) and Greg (path length) -- the last two are
    double backward through conv / upfirdn2d / bias_act / modulation -- against the reference's own
    SG2Loss / R1reg / PPLreg run on CPU (oracle/make_golden.py)."""
    from sgb200 import training
    z, meta = gold
    m = MODES[mode]
    torch.backends.cudnn.allow_tf32 = m['tf32']          # restored by conftest's autouse fixture
    cfg, _, _ = _build(meta, 'cpu', channels_last=channels_last, **m['over'])
    tr = training.Trainer(cfg, DEV)
    _load(z, tr.G, tr.D)
    zz = torch.from_numpy(z['z']).to(DEV)
    real = torch.from_numpy(z['real']).to(DEV)
    gains = meta['gains']

    def run(name, fn):
        mod = tr.G if name.startswith('G') else tr.D
        for p in mod.parameters():
            p.grad = None
        mod.requires_grad_(True)
        fn()
        mod.requires_grad_(False)
        return mod

    _check(z, 'Gmain', 'G.', run('Gmain', lambda: tr.phase_Gmain(zz, gains['Gmain'])), *phase_tolerances(mode, 'Gmain'))
    _check(z, 'Dmain', 'D.', run('Dmain', lambda: tr.phase_Dmain(zz, real, gains['Dmain'])), *phase_tolerances(mode, 'Dmain'))
    _check(z, 'Dreg', 'D.', run('Dreg', lambda: tr.phase_Dreg(real, gains['Dreg'])), *phase_tolerances(mode, 'Dreg'))
    pl_noise = torch.from_numpy(z['pl_noise']).to(DEV)
    tr.pl_mean.zero_()
    _check(z, 'Greg', 'G.', run('Greg', lambda: tr.phase_Greg(zz, gains['Greg'], pl_noise=pl_noise)), *phase_tolerances(mode, 'Greg'))


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['strict', 'tf32', 'fp16'])
def test_cuda_graph_phases_match_golden_and_eager(gold, mode):
    """The benched launch mode: every phase captured in a CUDA graph (TrainConfig.cuda_graphs) with z, the real images and the
    path-length noise as static inputs.  The gradients a replay leaves behind must match (a) the reference golden and (b)
    the same phase launched eagerly, and the warm-up that precedes the capture must not have advanced the training state."""
    from sgb200 import training
    z, meta = gold
    m = MODES[mode]
    torch.backends.cudnn.allow_tf32 = m['tf32']
    cfg, _, _ = _build(meta, 'cpu', channels_last=True, cuda_graphs=True, lr=0.0, force_flat_grads=(mode != 'strict'), **m['over'])     # lr 0: all four phases see
    tr = training.Trainer(cfg, DEV)                                                                # the golden weights
    _load(z, tr.G, tr.D)
    w0 = [p.detach().clone() for p in list(tr.G.parameters()) + list(tr.D.parameters())]
    zz = torch.from_numpy(z['z']).to(DEV)
    real = torch.from_numpy(z['real']).to(DEV)
    tr.static_z = {n: zz for n in ('Gmain', 'Greg', 'Dmain', 'Dreg')}
    tr.static_pl_noise = torch.from_numpy(z['pl_noise']).to(DEV)
    assert meta['gains'] == dict(Gmain=1, Dmain=1, Dreg=cfg.d_reg_interval, Greg=cfg.g_reg_interval)
    from sgb200.ops import upfirdn2d as _up
    _up.stats.update(sep=0, general=0, unseen_in_capture=0)
    out = tr.iteration(real, force_all_phases=True)          # builds the graphs (warm-up + capture) and replays all four
    print('upfirdn2d launches while building the graphs:', _up.stats)
    assert _up.stats['unseen_in_capture'] == 0, 'a FIR filter was first seen during graph capture (general kernel captured)'
    torch.cuda.synchronize()
    assert sorted(out) == ['Dmain', 'Dreg', 'Gmain', 'Greg'] and tr.replayed_launches > 0
    for a, b in zip(w0, list(tr.G.parameters()) + list(tr.D.parameters())):
        assert torch.equal(a, b.detach()), 'graph warm-up / lr=0 replay changed the weights'
    assert float(tr.pl_mean) != 0.0
    # Gmain / Greg (and Dmain / Dreg) share one flat gradient buffer per module: replay phase by phase and check each
    graph_grads = {}
    for name in ('Gmain', 'Dmain', 'Dreg', 'Greg'):
        mod = tr.G if name.startswith('G') else tr.D
        g1, g2, val, launches, grads = tr._graphs['graphs'][name]
        assert g2 is None
        tr.pl_mean.zero_()                                    # the golden path-length run starts from pl_mean = 0
        g1.replay()
        torch.cuda.synchronize()
        assert torch.isfinite(val).all()
        graph_grads[name] = [None if g is None else g.detach().clone() for g in grads]
        for p, g in zip(mod.parameters(), grads):
            p.grad = g
        _check(z, name, name[0] + '.', mod, *phase_tolerances(mode, name))
    # (b) eager launches of the same phases on the same state
    for ph in tr.phases:
        tr.pl_mean.zero_()
        tr._phase_grads(ph, real, zz)
        eager = [None if p.grad is None else p.grad.detach().clone() for p in ph['module'].parameters()]
        floor = 1e-2 * max(float(ge.abs().max()) for ge in eager if ge is not None)       # as in helpers.check_phase_grads
        for ge, gg in zip(eager, graph_grads[ph['name']]):
            assert (ge is None) == (gg is None)
            if ge is None:
                continue
            scale = max(float(ge.abs().max()), floor)
            # same kernels, same inputs; the weight-gradient kernels reduce with atomics, so the sums differ in the last bits
            # (more visibly in the second-order phases, whose operands are themselves such sums)
            rtol = 1e-4 if mode == 'strict' else 5e-3
            assert float((ge - gg).abs().max()) <= rtol * scale + 1e-12, f"{ph['name']}: graph replay differs from eager launch"


@pytest.mark.gpu
def test_fp16_network_against_fp32_oracle():
    """num_fp16_res=2 + conv_clamp=256 (the config-f / sg2attent numerics) at 32x32: 1e-2 against the fp32 oracle
    run with the same weights."""
    from sgb200 import training
    cfg = training.TrainConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=1024, channel_max=64, map_layers=2,
                               num_fp16_res=2, conv_clamp=256.0, batch_gpu=4, style_mixing_prob=0.0, noise_mode='const',
                               use_ema=False, mbstd_group_size=2)
    torch.manual_seed(0)
    G, D = training.build_networks(cfg, DEV)
    with torch.no_grad():
        for n_, p in G.named_parameters():
            if n_.endswith('noise_strength'):
                p.fill_(0.1)
    ocfg = RN.NetConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=1024, channel_max=64, map_layers=2,
                        num_fp16_res=0, conv_clamp=256.0, d_arch='resnet', mbstd_group_size=2)
    GP = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    DP = {k: v.detach().cpu() for k, v in D.state_dict().items()}
    zz = torch.randn(4, 64)
    with torch.no_grad():
        img = G.synthesis(G.mapping(zz.to(DEV), None, skip_w_avg_update=True), noise_mode='const')
        logits = D(img, None)
        img_o = RN.g_synthesis(GP, RN.g_mapping(GP, zz, ocfg), ocfg, noise='const')
        logits_o = RN.d_forward(DP, img.cpu().float(), ocfg)
    assert img.dtype == torch.float32
    assert_close(img, img_o, 1e-2, 'img fp16 blocks')
    assert_close(logits, logits_o, 2e-2, 'logits fp16 blocks')


@pytest.mark.gpu
def test_trainer_iteration_runs_all_phases():
    from sgb200 import training, _lib
    cfg = training.TrainConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=512, channel_max=32, map_layers=2,
                               batch_gpu=4, mbstd_group_size=2, g_reg_interval=4, d_reg_interval=2)
    tr = training.Trainer(cfg, DEV)
    real = torch.randint(0, 256, [4, 3, 32, 32], dtype=torch.uint8, device=DEV)
    before = [p.detach().clone() for p in tr.G.parameters()]
    n0 = _lib.launch_count()
    seen = set()
    for i in range(4):
        out = tr.iteration(real)
        seen |= set(out)
        for v in out.values():
            assert torch.isfinite(v).all()
    assert seen == {'Gmain', 'Greg', 'Dmain', 'Dreg'}
    assert _lib.launch_count() > n0
    assert any((a != b.detach()).any() for a, b in zip(before, tr.G.parameters()))


@pytest.mark.gpu
def test_cuda_graph_training_iterations():
    """Each phase captured in a CUDA graph and replayed: losses stay finite, parameters move, the lazy
    regularisation schedule is honoured, and the replayed launch count is what the capture recorded."""
    import torch
    from sgb200 import training
    dev = torch.device('cuda', 0)
    cfg = training.TrainConfig(img_resolution=32, z_dim=64, w_dim=64, channel_base=1024, channel_max=64, map_layers=2,
                               batch_gpu=4, mbstd_group_size=2, g_reg_interval=4, d_reg_interval=2, cuda_graphs=True)
    tr = training.Trainer(cfg, dev)
    real = torch.randint(0, 256, [4, 3, 32, 32], dtype=torch.uint8, device=dev)
    p0 = [p.detach().clone() for p in tr.G.parameters()]
    seen = []
    for i in range(4):
        out = tr.iteration(real)
        torch.cuda.synchronize()
        seen.append(sorted(out))
        assert all(torch.isfinite(v).all() for v in out.values()), (i, out)
    assert seen[0] == ['Dmain', 'Dreg', 'Gmain', 'Greg'] and seen[1] == ['Dmain', 'Gmain'] and seen[2] == ['Dmain', 'Dreg', 'Gmain']
    assert tr.replayed_launches > 0
    assert any((a - b.detach()).abs().max() > 0 for a, b in zip(p0, tr.G.parameters()))
    assert all(torch.isfinite(p).all() for p in list(tr.G.parameters()) + list(tr.D.parameters()))


@pytest.mark.gpu
def test_multi_tensor_nan_to_num():
    """sgb_nan_to_num_multi == torch.nan_to_num per tensor (trainers.py:745-748), > 96 tensors, empty and odd sizes."""
    from sgb200.training import nan_to_num_
    torch.manual_seed(3)
    sizes = [0, 1, 3, 7, 512, 1000, 4097, 512 * 9 * 31] + [5 + i for i in range(120)]
    ts = []
    for i, n in enumerate(sizes):
        t = torch.randn(n, device='cuda') * 1e3
        if n > 0:
            t[torch.randint(0, n, [max(1, n // 7)], device='cuda')] = float('nan')
            t[torch.randint(0, n, [max(1, n // 9)], device='cuda')] = float('inf')
            t[torch.randint(0, n, [max(1, n // 11)], device='cuda')] = float('-inf')
        ts.append(t)
    ts.append(torch.randn(6, 5, device='cuda').t())                      # non-contiguous: falls back per tensor
    ts.append(torch.randn(9, device='cuda', dtype=torch.float16))
    ts[-1][2] = float('nan')
    want = [torch.nan_to_num(t, nan=0, posinf=1e5, neginf=-1e5) for t in ts]
    nan_to_num_(ts, nan=0, posinf=1e5, neginf=-1e5)
    for a, b in zip(ts, want):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['strict', 'tf32'])
def test_config_A_real_size_against_oracle(mode):
    """BASELINE configs[0] at its real size (sg2ada.yaml, 64x64, batch 8, 512-channel layers, D 'orig', mbstd over the whole
    batch): G forward, D forward and the parameter gradients of Gmain / Dmain / Dreg (R1) on the GPU against the CPU
    oracle (oracle/ref_networks.py, pinned to the reference by net_tiny.npz) with the same weights and inputs."""
    from sgb200 import training
    tf32 = mode == 'tf32'
    # (median over tensors, worst tensor).  Calibration (profiles/r2_parity_report.md, the reference's own GPU path against its
    # own CPU run at this size): strict fp32 cuDNN median 1.0e-4 ... 5.4e-4, worst 5.6e-3; cuDNN TF32 median up to 3.2e-2,
    # worst 5.2e-1 (noise_strength scalars: sums with heavy cancellation).  The bounds below are ~2x the strict figures
    # and ~1.2x the TF32 ones.
    tol_med, tol_worst = (4e-2, 6e-1) if tf32 else (1e-3, 1e-2)
    torch.backends.cudnn.allow_tf32 = tf32
    cfg = training.config_sg2ada64(noise_mode='const', use_ema=False)
    torch.manual_seed(0)
    tr = training.Trainer(cfg, DEV)
    with torch.no_grad():
        for n_, p in tr.G.named_parameters():
            if n_.endswith('noise_strength'):
                p.fill_(0.1)
    ocfg = RN.NetConfig(img_resolution=64, z_dim=512, w_dim=512, channel_base=32768, channel_max=512, map_layers=2, num_fp16_res=0,
                        conv_clamp=None, d_arch='orig', mbstd_group_size=32)
    # the oracle runs in float64: at 512 channels an fp32 CPU run carries as much summation noise as the kernels under test
    GP = {k: v.detach().cpu().double() for k, v in tr.G.state_dict().items()}
    DP = {k: v.detach().cpu().double() for k, v in tr.D.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    zz = torch.randn(8, 512, generator=g).double()
    real = (torch.rand(8, 3, 64, 64, generator=g) * 2 - 1).double()
    with torch.no_grad():
        img = tr.G.synthesis(tr.G.mapping(zz.float().to(DEV), None, skip_w_avg_update=True), noise_mode='const')
        logits = tr.D(img, None)
        img_o = RN.g_synthesis(GP, RN.g_mapping(GP, zz, ocfg), ocfg, noise='const')
        logits_o = RN.d_forward(DP, img_o, ocfg)
    assert_close(img, img_o, 1e-2 if tf32 else 1e-4, 'img 64x64')
    assert_close(logits, logits_o, 2e-2 if tf32 else 2e-4, 'logits 64x64')

    def grads_gpu(name, fn):
        mod = tr.G if name.startswith('G') else tr.D
        for p in mod.parameters():
            p.grad = None
        mod.requires_grad_(True)
        fn()
        mod.requires_grad_(False)
        return {k: p.grad.detach().cpu() for k, p in mod.named_parameters() if p.grad is not None}

    def compare(name, got, want):
        floor = 1e-2 * max(float(v.abs().max()) for v in want.values())
        assert set(got) == set(want), (name, set(got) ^ set(want))
        errs = sorted((float((got[k].double() - want[k].double()).abs().max()) / max(float(want[k].abs().max()), floor), k) for k in want)
        worst, wk = errs[-1]
        median = errs[len(errs) // 2][0]
        print(f'config A [{mode}] {name}: worst {worst:.2e} ({wk}), median {median:.2e}')
        assert median <= tol_med, f'{name}: median rel err {median:.3e} > {tol_med:.1e}'
        assert worst <= tol_worst, f'{name} {wk}: rel err {worst:.3e} > {tol_worst:.1e}'

    go = RN.phase_gmain(GP, DP, zz, ocfg, ocfg, noise='const')[1]
    compare('Gmain', grads_gpu('Gmain', lambda: tr.phase_Gmain(zz.float().to(DEV), 1)), go)
    do = RN.phase_dmain(GP, DP, zz, real, ocfg, ocfg, noise='const')[1]
    compare('Dmain', grads_gpu('Dmain', lambda: tr.phase_Dmain(zz.float().to(DEV), real.float().to(DEV), 1)), do)
    ro = RN.phase_dreg(DP, real, ocfg, r1_gamma=cfg.r1_gamma, gain=4)[1]
    compare('Dreg', grads_gpu('Dreg', lambda: tr.phase_Dreg(real.float().to(DEV), 4)), ro)
