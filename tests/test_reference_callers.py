"""Drop-in check: the reference's own smoke test ``tests/test_pipelines.py:9-63`` (``test_chattts_plus_pipeline``) runs UNCHANGED
against this package — ``chattts_plus.pipelines.ChatTTSPlusPipeline(OmegaConf.load("configs/infer/chattts_plus.yaml"), device=cuda)``,
``pipeline.infer(..., lora_path=..., speaker_emb_path='')`` — with a synthetic ``CHATTTS_PLUS_CHECKPOINT_DIR`` (real shapes, the files
the YAML names, a small pickled ``BertTokenizerFast``, a peft-format LoRA directory at the path the script hard-codes).

The script is executed from ``oracle/_ref/test_pipelines.py``, an untracked copy made by ``oracle/fetch_ref.py`` (run by
``__graft_entry__.build()`` wherever the reference checkout exists); the test is skipped when that copy is absent.
"""
import glob
import os
import shutil
import subprocess
import sys
import wave

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF_TEST = os.path.join(ROOT, "oracle", "_ref", "test_pipelines.py")
TEXT = "我们针对对话式任务进行了优化，能够实现自然且富有表现力的合成语音"          # the sentence the reference script synthesises
LORA_REL = "outputs/leijun_lora-1734532984.1128285/checkpoints/step-2000"    # the adapter path the reference script hard-codes


@pytest.mark.gpu
def test_reference_test_pipelines_runs_unchanged(tmp_path):
    if not os.path.exists(REF_TEST):
        pytest.skip("oracle/_ref/test_pipelines.py is absent (python oracle/fetch_ref.py copies it where /root/reference exists)")
    from chatttsplus_b200 import synth
    proj = tmp_path / "project"
    ckpt = tmp_path / "checkpoints"
    (proj / "configs" / "infer").mkdir(parents=True)
    shutil.copy(os.path.join(ROOT, "configs", "infer", "chattts_plus.yaml"), proj / "configs" / "infer" / "chattts_plus.yaml")
    synth.write_checkpoint_dir(str(ckpt), texts=[TEXT], lora_dir=str(proj / LORA_REL))
    env = dict(os.environ, CHATTTS_PLUS_CHECKPOINT_DIR=str(ckpt), CHATTTS_PLUS_PROJECT_DIR=str(proj), CHATTTS_PLUS_LOG_DIR=str(tmp_path / "logs"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_caller.py"), REF_TEST, "test_chattts_plus_pipeline"],
                       cwd=str(proj), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "total infer time" in r.stdout and "load lora into gpt" in r.stdout and "unload lora" in r.stdout, r.stdout[-2000:]
    wavs = glob.glob(str(proj / "results" / "chattts_plus" / "*.wav"))
    assert len(wavs) == 1, wavs
    with wave.open(wavs[0], "rb") as f:
        assert f.getframerate() == 24000 and f.getnframes() > 24000 // 4
    assert glob.glob(str(proj / "results" / "speakers" / "*.pt")), "the random speaker is saved like chattts_plus_pipeline.py:548-557"
