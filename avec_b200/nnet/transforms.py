"""On-device input augmentation with the reference's training recipe:

  VideoAugment     training_video_transform of configs/LRS23/{VO,AV}/EffConfInterCTC.py:60-70 / 82-88 - torchvision RandomCrop(88, 88),
                   RandomHorizontalFlip, nnet.TimeMaskSecond(T_second=0.4, num_mask_second=1.0, fps=25, mean_frame=True)
                   (nnet/transforms.py:108-126) - applied per sample to the whole padded batch by two kernels (csrc/train.cu), no Python
                   loop over samples or masks, no host sync; in eval() it is the configs' evaluation transform, CenterCrop.
  align_video_to_audio   nnet/transforms.py:169-180 (Tv = Ta // 640 + 1).
"""
import torch
import torch.nn as nn

from .. import ops


class VideoAugment(nn.Module):
    def __init__(self, crop_size=(88, 88), flip_p=0.5, T_second=0.4, num_mask_second=1.0, fps=25.0):
        super().__init__()
        self.crop_size, self.flip_p = tuple(crop_size), flip_p
        self.mask_T, self.num_mask_second, self.fps = int(T_second * fps), num_mask_second, fps      # TimeMaskSecond.__init__

    def forward(self, video, lengths=None):
        """video (B, T, Hi, Wi[, 1]) fp32 on the device, lengths (B,) valid frames -> (B, T, Ho, Wo[, 1])"""
        last1 = video.dim() == 5
        v = video[..., 0] if last1 else video
        Ho, Wo = self.crop_size
        if self.training:
            ln = lengths.to(device=v.device, dtype=torch.long) if lengths is not None else None
            out = ops.video_augment(v.float().contiguous(), ln, ops.RNG.next_site(), self.crop_size, self.flip_p, self.mask_T, self.fps,
                                    self.num_mask_second)
        else:
            oy, ox = int(round((v.shape[2] - Ho) / 2.0)), int(round((v.shape[3] - Wo) / 2.0))     # torchvision CenterCrop
            out = v[:, :, oy:oy + Ho, ox:ox + Wo].contiguous()
        return out.unsqueeze(-1) if last1 else out

    def extra_repr(self):
        return f"crop={self.crop_size}, flip_p={self.flip_p}, mask_T={self.mask_T}, num_mask_second={self.num_mask_second}, fps={self.fps}"


def align_video_to_audio(video, audio):
    """video (T, H, W, C), audio (L,): pad / cut the video to L // 640 + 1 frames (nnet/transforms.py:169-180)"""
    Tv = audio.shape[-1] // 640 + 1
    if video.shape[0] < Tv:
        return torch.nn.functional.pad(video, (0, 0, 0, 0, 0, 0, 0, Tv - video.shape[0]))
    return video[:Tv]
