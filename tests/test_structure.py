"""CPU tests of the host-side mirror: state_dict keys / shapes / parameter counts equal the reference's (known answers
from demo.ipynb cell 12), length bookkeeping, mel filterbank and positional table restatements."""
import json
import os

import pytest
import torch

from avec_b200 import nnet
from avec_b200.nnet.modules import rel_pos_table
from conftest import GOLDEN
from oracle import restate


@pytest.mark.parametrize("name", ["AudioEfficientConformerInterCTC", "VisualEfficientConformerInterCTC", "AudioVisualEfficientConformerInterCTC"])
def test_state_dict_contract(name):
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))[name]
    m = getattr(nnet, name)()
    sd = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert list(sd.keys()) == list(ref["keys"].keys())
    assert sd == ref["keys"]
    assert m.num_params() == ref["params"]


def test_known_parameter_counts():
    assert nnet.VisualEfficientConformerInterCTC().num_params() == 40903112
    assert nnet.AudioVisualEfficientConformerInterCTC().num_params() == 61738836
    assert nnet.AudioEfficientConformerInterCTC(interctc_blocks=[]).num_params() == 31562460


def test_mel_filterbank_matches_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    fb = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 80, 16000, None, "htk")
    assert torch.equal(fb, nnet.mel_filterbank())
    assert torch.equal(fb, restate.mel_filterbank())


def test_rel_pos_table():
    T, D = 7, 12
    pe = rel_pos_table(T, D, "cpu", torch.float32)
    assert torch.equal(pe, restate.rel_pos_table(T, D))
    # row r is the sinusoid of relative position T-1-r
    r, k = 2, 3
    pos = T - 1 - r
    assert abs(pe[r, 2 * k].item() - torch.sin(torch.tensor(pos / 10000 ** (2 * k / D))).item()) < 1e-6


def test_length_bookkeeping():
    L0 = torch.tensor([64000, 63999, 640, 1])
    f = torch.div(L0, 160, rounding_mode="floor") + 1
    assert f.tolist() == [401, 400, 5, 1]
    s = torch.div(f - 1, 2, rounding_mode="floor") + 1
    assert s.tolist() == [201, 200, 3, 1]


def test_no_cpu_fallback():
    """the product path has no CPU / eager fallback: CPU tensors are rejected loudly"""
    m = nnet.AudioEfficientConformerInterCTC()
    m.train()
    with pytest.raises(RuntimeError, match="CUDA"):
        m((torch.zeros(1, 3200), torch.tensor([3200])))


def test_training_config_matches_reference():
    """train() runs the reference's training graph: dropout 0.1 at the six sites of every block + the stack input, SpecAugment
    (2, 27, 5, 0.05) in the audio encoder (networks.py:327,347-353); zero_dropout() gives the deterministic parity graph"""
    m = nnet.AudioVisualEfficientConformerInterCTC()
    drops = [d for d in m.modules() if isinstance(d, torch.nn.Dropout)]
    assert len(drops) == 24 * 6 + 3 and all(d.p == 0.1 for d in drops)
    sa = m.encoder.audio_encoder.spec_augment
    assert sa.params() == (2, 27, 5, 0.05) and sa.enabled
    nnet.zero_dropout(m)
    assert all(d.p == 0.0 for d in drops) and not sa.enabled
