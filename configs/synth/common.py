"""Synthetic twins of the reference's LRS2/3 configs (configs/LRS23/{AO,VO,AV}/*.py): the same module-level names main.py reads
(main.py:49,66-89) - `model`, `training_dataset`, `evaluation_dataset`, `precision`, `callback_path`, `epochs`, ... - with a
seeded in-memory dataset in place of LRS2/3 (no network, no checkpoints).  The dataset yields the reference's sample tuple
(video [T,H,W,1], audio [L], label, video_len, audio_len, label_len), nnet/datasets.py:326-366, and uses the reference's own
nnet.CollateFn with the axis mapping of the real configs.

Run from the repository root with the reference tree on sys.path (its main.py's directory):

    python baseline/_ref/main.py -c configs/synth/AV.py --steps_per_epoch 2          # or /root/reference/main.py in the container
"""
import os

import torch

import nnet            # the reference's package (main.py's directory is sys.path[0])
import avec_b200

avec_b200.patch_reference(nnet)


class SyntheticAV(nnet.datasets.Dataset):
    """seeded synthetic utterances: SECONDS of 16 kHz audio, Tv = L // 640 + 1 video frames of 88x88 (the align rule of
    nnet/transforms.py:169-180), 12-token labels, lengths ragged over the dataset"""

    def __init__(self, batch_size, collate_fn, n=16, seconds=1.0, seed=0, shuffle=False):
        super().__init__(batch_size=batch_size, collate_fn=collate_fn, root=None, shuffle=shuffle)
        self.n, self.L, self.seed = n, int(16000 * seconds), seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000 + i)
        L = self.L - 640 * (i % 3)
        Tv = L // 640 + 1
        audio = 0.1 * torch.randn(L, generator=g)
        video = torch.randn(Tv, 88, 88, 1, generator=g).clamp_(-1, 1)
        label = torch.randint(1, 256, (12,), generator=g)
        return video, audio, label, torch.tensor(Tv), torch.tensor(L), torch.tensor(12)


precision = torch.bfloat16
epochs = 1
eval_training = False
saving_period_epoch = 1
callback_root = os.environ.get("AVEC_SYNTH_CALLBACKS", "gpurun_out/callbacks_synth")
