#!/usr/bin/env bash
# Stage the reference's OWN sources for the hot path and its caller (BASELINE configs[2]) under oracle/_ref/ so that they
# can travel to the GPU box with `gpurun` (oracle/_ref/ is git-ignored: nothing of the reference enters the history).
# The reference is pure Python, so "building" it is copying the files it needs, unmodified:
#   losses/                 the spectral losses themselves (stft_loss.py, mel_loss.py) + the other criteria
#   trainer/                trainer.denoise.Trainer._train_step / TrainerGAN._metric_loss -- the caller of the hot path
#   models/autoencoder/, models/utils.py, layers/   the symAD generator the denoise trainer drives
#   models/vocoder/         HiFiGAN generator + HiFiGAN / UnivNet discriminators (autoencoder and vocoder trainers)
#   config/denoise/symAD_vctk_48000_hop300.yaml     generator / loss / optimizer hyper-parameters of configs[2]
#   config/autoencoder/*.yaml, config/vocoder/*.yaml  the shipped hyper-parameters of the other two trainers
# Used ONLY by tests/, bench.py's reference arm / cpu_baseline and profiles/trainer_step.py (see oracle/ref_loader.py).
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/losses/stft_loss.py" ]; then
  echo "make_ref: reference not found at $SRC (nothing staged)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/models" "$DST/config/denoise" "$DST/config/autoencoder" "$DST/config/vocoder"
cp -r "$SRC/losses" "$DST/losses"
cp -r "$SRC/trainer" "$DST/trainer"
cp -r "$SRC/layers" "$DST/layers"
cp -r "$SRC/models/autoencoder" "$DST/models/autoencoder"
cp "$SRC/models/utils.py" "$DST/models/utils.py"
cp -r "$SRC/models/vocoder" "$DST/models/vocoder"
cp "$SRC/config/denoise/symAD_vctk_48000_hop300.yaml" "$DST/config/denoise/"
cp "$SRC"/config/autoencoder/*.yaml "$DST/config/autoencoder/"
cp "$SRC"/config/vocoder/*.yaml "$DST/config/vocoder/"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$DST" && find . -type f | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "make_ref: staged $(find "$DST" -name '*.py' | wc -l) python files + $(find "$DST" -name '*.yaml' | wc -l) yaml under $DST"
