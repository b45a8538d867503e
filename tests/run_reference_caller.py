"""Child process of tests/test_reference_callers.py: executes ONE function of the reference's own tests/test_pipelines.py, unchanged,
from the untracked copy under oracle/_ref/ (see oracle/fetch_ref.py).  The environment (CHATTTS_PLUS_CHECKPOINT_DIR,
CHATTTS_PLUS_PROJECT_DIR) and the working directory are prepared by the parent; this script only adds the repository root to
sys.path (so ``chattts_plus`` / ``omegaconf`` resolve to the drop-in shims) and, when torchaudio's file writer has no backend in
the image (torchcodec absent), installs a PCM-16 WAV writer behind ``torchaudio.save`` — an environment accommodation, the
reference script itself is not edited.

    python tests/run_reference_caller.py <path to test_pipelines.py> <function name>
"""
import importlib.util
import os
import sys
import wave

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def _install_wav_writer():
    import numpy as np
    import torch
    import torchaudio
    try:
        import torchcodec  # noqa: F401
        return
    except Exception:
        pass

    def save(path, src, sample_rate, **kw):
        a = (src.detach().cpu().float().clamp(-1, 1) * 32767.0).round().to(torch.int16).numpy()
        with wave.open(str(path), "wb") as f:
            f.setnchannels(a.shape[0])
            f.setsampwidth(2)
            f.setframerate(int(sample_rate))
            f.writeframes(np.ascontiguousarray(a.T).tobytes())
    torchaudio.save = save


def main():
    path, fn = sys.argv[1], sys.argv[2]
    _install_wav_writer()
    spec = importlib.util.spec_from_file_location("reference_test_pipelines", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    getattr(mod, fn)()


if __name__ == "__main__":
    main()
