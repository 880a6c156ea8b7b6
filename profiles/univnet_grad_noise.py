"""Why gradients through the reference's UnivNet / HiFiGAN discriminators cannot be compared at 1e-4 between two fp32 front-ends:
LeakyReLU sign flips make the whole-network gradient differ discretely (the period discriminators alone, with no spectrogram in
them, are 2e-3 from their own fp64 evaluation).  Output: profiles/r4e_dbg_univ.txt.  Measurement script, not product."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchaudio
import dl_speech_enhancement_b200 as pkg
from oracle import trainer_harness as th
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
x0 = 0.1*torch.randn(4,1,9600, device=dev)
for (n,h,w) in ((1024,120,600),(2048,240,1200),(512,50,240)):
    win = torch.hann_window(w, device=dev)
    def run(fn, x, dt):
        x = x.to(dt).clone().requires_grad_(True)
        m = fn(x, pad=w//2, window=win.to(dt), n_fft=n, hop_length=h, win_length=w, power=1.0, normalized=False)
        return m, x
    m_ref, xr = run(torchaudio.functional.spectrogram, x0, torch.float32)
    g = torch.randn_like(m_ref)
    for name, gg in (("random", g), ("ones", torch.ones_like(g)), ("inv_mag", 1.0/(m_ref.detach()+1e-3))):
        m_ref, xr = run(torchaudio.functional.spectrogram, x0, torch.float32); m_ref.backward(gg)
        m64, x64 = run(torchaudio.functional.spectrogram, x0, torch.float64); m64.backward(gg.double())
        m_o, xo = run(pkg.spectrogram, x0, torch.float32); m_o.backward(gg)
        e = lambda a: float((a.double()-x64.grad).norm()/x64.grad.norm())
        print(n, name, "fwd", float((m_o-m_ref).norm()/m_ref.norm()), "grad ours", e(xo.grad), "ref32", e(xr.grad))
ns = th.load()
cfg = ns.configs["vocoder/AudioDec_v3_symADuniv_vctk_48000_hop300_clean"]
torch.manual_seed(3)
disc = ns.UnivNetDiscriminator(**cfg["discriminator_params"]).to(dev)
disc64 = ns.UnivNetDiscriminator(**cfg["discriminator_params"]).to(dev); disc64.load_state_dict(disc.state_dict()); disc64 = disc64.double()
adv = ns.GeneratorAdversarialLoss(**cfg["generator_adv_loss_params"]).to(dev)
mod = ns.discriminator_module
stock = mod.spectrogram
def run(fe, net, x, part):
    mod.spectrogram = fe
    try:
        x = x.clone().requires_grad_(True); o = getattr(net, part)(x) if part else net(x); adv(o).backward(); return o, x.grad
    finally:
        mod.spectrogram = stock
for part in ("mrsd", "mpd", None):
    o_ref, g_ref = run(stock, disc, x0, part)
    o_ref2, g_ref2 = run(stock, disc, x0, part)
    o_our, g_our = run(pkg.spectrogram, disc, x0, part)
    o64, g64 = run(stock, disc64, x0.double(), part)
    e = lambda a: float((a.double()-g64).norm()/g64.norm())
    print(part, ": ours", e(g_our), "ref32", e(g_ref), "ref32 rerun", e(g_ref2), "|g64|", float(g64.norm()))
    d = (g_our.double()-g64).abs(); d2 = (g_ref.double()-g64).abs()
    print("  ours err by position:", ["%.1e" % float(d[:,0,i:i+600].max()) for i in range(0,9600,600)], "gmax %.2e" % float(g64.abs().max()))
    print("  ref32 err by position:", ["%.1e" % float(d2[:,0,i:i+600].max()) for i in range(0,9600,600)])
# per-discriminator of the mrsd
for i, dsc in enumerate(disc.mrsd.discriminators):
    def runi(fe, net, x):
        mod.spectrogram = fe
        try:
            x = x.clone().requires_grad_(True); o = net(x); adv([o]).backward(); return x.grad
        finally:
            mod.spectrogram = stock
    g_ref = runi(stock, dsc, x0); g_our = runi(pkg.spectrogram, dsc, x0); g64 = runi(stock, disc64.mrsd.discriminators[i], x0.double())
    e = lambda a: float((a.double()-g64).norm()/g64.norm())
    print("mrsd", i, dsc.fft_size, dsc.hop_size, dsc.win_length, "ours", e(g_our), "ref32", e(g_ref))
