import cProfile, pstats, sys, os, torch, time
sys.path.insert(0, os.getcwd())
import dl_speech_enhancement_b200 as pkg
dev = torch.device("cuda:0")
MEL_KW = dict(fs=48000, fft_sizes=[2048], hop_sizes=[300], win_lengths=[None], window="hann_window", num_mels=80, fmin=0, fmax=24000, log_base=None)
stft = pkg.MultiResolutionSTFTLoss().to(dev); mel = pkg.MultiMelSpectrogramLoss(**MEL_KW).to(dev)
y = 0.1 * torch.randn(16, 1, 48000, device=dev); x = (y + 0.05 * torch.randn_like(y)).requires_grad_(True)
def step():
    x.grad = None
    ml = mel(x, y); sc, mag = stft(x, y)
    (sc + mag + ml).backward()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print("host issue time per step %.1f us, with sync %.1f us" % (t_issue / 300 * 1e6, (time.perf_counter() - t0) / 300 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
