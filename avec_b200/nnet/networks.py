"""Encoders of the Efficient Conformer family (reference nnet/networks.py:32-146, 202-579) on the fused sm_100a
kernels.  Constructor arguments, attribute names and state_dict keys follow the reference so released checkpoints load
(SURVEY A.2); internally activations are channels-last in the compute dtype (bf16 production / fp32 parity)."""
import os

import torch
import torch.nn as nn

from .. import functional as AF
from .. import ops
from .blocks import ConformerBlock, ResNetBlock
from .layers import Linear, Conv2d, Conv3d, Placeholder, Dropout, Swish, mel_filterbank
from .modules import InterCTCResModule, FusionModule
from .preprocessing import SpecAugment


class ConformerInterCTC(nn.Module):
    def __init__(self, dim_model, num_blocks, interctc_blocks, vocab_size, loss_prefix="ctc", att_params=None, conv_params=None,
                 ff_ratio=4, drop_rate=0.1, pos_embedding=None, mask=None, conv_stride=1, batch_norm=True):
        super().__init__()
        assert pos_embedding is None
        self.interctc_blocks = interctc_blocks
        self.loss_prefix = loss_prefix
        if isinstance(dim_model, int):
            dim_model = [dim_model]
        if isinstance(num_blocks, int):
            num_blocks = [num_blocks]
        self.dropout = Dropout(p=drop_rate)
        i = 1
        self.conformer_blocks = nn.ModuleList()
        self.interctc_modules = nn.ModuleList()
        for stage_id in range(len(num_blocks)):
            for block_id in range(num_blocks[stage_id]):
                down_block = (block_id == num_blocks[stage_id] - 1) and (stage_id < len(num_blocks) - 1)
                self.conformer_blocks.append(ConformerBlock(
                    dim_model=dim_model[stage_id],
                    dim_expand=dim_model[stage_id + (1 if down_block else 0)],
                    ff_ratio=ff_ratio,
                    drop_rate=drop_rate,
                    att_params=att_params[stage_id] if isinstance(att_params, list) else att_params,
                    conv_stride=1 if not down_block else (conv_stride[stage_id] if isinstance(conv_stride, list) else conv_stride),
                    conv_params=conv_params[stage_id] if isinstance(conv_params, list) else conv_params,
                    batch_norm=batch_norm))
                if i in interctc_blocks:
                    self.interctc_modules.append(InterCTCResModule(dim_model[stage_id + (1 if down_block else 0)], vocab_size))
                i += 1

    def forward(self, x, lengths):
        """x (B,T,D) compute dtype, lengths (B,) int -> (x, lengths, {prefix_i: [logits fp32, lengths]})."""
        klen = lengths.to(device=x.device, dtype=torch.int32) if lengths is not None else None
        x = self.dropout(x)                                                  # networks.py:268-269
        interctc_outputs = {}
        j = 0
        for i, block in enumerate(self.conformer_blocks):
            x = block(x, klen=klen)
            logits = None
            if i + 1 in self.interctc_blocks:
                x, logits = self.interctc_modules[j](x)
                j += 1
                key = self.loss_prefix + "_" + str(i)
            if block.stride > 1:
                # mask[:, :, ::s, ::s] keeps key j' iff j'*s < len  <=>  j' < (len-1)//s + 1 (networks.py:294-302)
                if lengths is not None:
                    lengths = torch.div(lengths - 1, block.stride, rounding_mode="floor") + 1
                    klen = lengths.to(device=x.device, dtype=torch.int32)
            if logits is not None:
                interctc_outputs[key] = [logits, lengths]
        return x, lengths, interctc_outputs


def _att(cls, heads, **kw):
    p = {"num_heads": heads, "attn_drop_rate": 0.0, "num_pos_embeddings": 10000, "weight_init": "default", "bias_init": "default"}
    p.update(kw)
    return {"class": cls, "params": p}


class _AudioPreprocessingParams(nn.Module):
    """Holds the two persistent torchaudio buffers of the reference (…Spectrogram.window, …MelScale.fb)."""

    def __init__(self):
        super().__init__()
        self.Spectrogram = nn.Module()
        self.Spectrogram.register_buffer("window", torch.hann_window(400))
        self.MelScale = nn.Module()
        self.MelScale.register_buffer("fb", mel_filterbank())


class _SubsamplingParams(nn.Module):
    """ConvNeuralNetwork(dim_input=1, dim_layers=C, kernel 3, stride 2, BatchNorm2d, Swish) parameter tree (modules.py:70-130)."""

    def __init__(self, filters):
        super().__init__()
        self.layers = nn.ModuleList([nn.Sequential(
            Conv2d(1, filters, 3, stride=2), nn.BatchNorm2d(filters), Swish(), Placeholder("Identity"))])


class AudioEfficientConformerEncoder(nn.Module):
    def __init__(self, include_head=True, vocab_size=256, att_type="patch", interctc_blocks=[3, 6, 10, 13], num_blocks=[5, 6, 5],
                 loss_prefix="ctc"):
        super().__init__()
        assert att_type in ["regular", "grouped", "patch"]
        filters, n_mels, dim_model, heads = 180, 80, [180, 256, 360], 4
        self.audio_preprocessing = _AudioPreprocessingParams()
        self.spec_augment = SpecAugment(mF=2, F=27, mT=5, pS=0.05)          # networks.py:347-353
        self.unsqueeze = Placeholder("Unsqueeze")
        self.subsampling_module = _SubsamplingParams(filters)
        self.reshape = Placeholder("Reshape")
        self.transpose = Placeholder("Transpose")
        self.linear = Linear(filters * n_mels // 2, dim_model[0])
        reg = _att("RelPos1dMultiHeadAttention", heads)
        if att_type == "grouped":   # Efficient-Conformer grouped attention (group 3 in stage 1), networks.py:389-393
            def grp(g):
                return {"class": "GroupedRelPosMultiHeadSelfAttention",
                        "params": {"num_heads": heads, "group_size": g, "attn_drop_rate": 0.0, "max_pos_encoding": 10000, "causal": False}}
            att_list = [grp(3), grp(1), grp(1)]
        else:
            first = reg if att_type == "regular" else _att("RelPosPatch1dMultiHeadAttention", heads, patch_size=3)
            att_list = [first, reg, reg]
        self.back_end = ConformerInterCTC(
            dim_model=dim_model, num_blocks=num_blocks, interctc_blocks=interctc_blocks, vocab_size=vocab_size,
            att_params=att_list, conv_params={"class": "Conv1d", "params": {"padding": "same", "kernel_size": 15}},
            ff_ratio=4, drop_rate=0.1, conv_stride=2, batch_norm=True, loss_prefix=loss_prefix)
        self.head = Linear(dim_model[-1], vocab_size) if include_head else nn.Identity()
        self._filters, self._nf = filters, n_mels // 2
        # kernel layout of the 7200->180 projection: the fused stem emits features as (f, c), the checkpoint stores (c, f)
        C, Fq = filters, n_mels // 2
        self._proj_layout = (
            "stem_proj",
            lambda w: w.view(w.shape[0], C, Fq).permute(0, 2, 1),          # a VIEW: rows of the [N, Fq*C] operand (avec_b200.weights)
            lambda dw: dw.view(dw.shape[0], Fq, C).permute(0, 2, 1).reshape(dw.shape[0], C * Fq),
        )

    def forward(self, x, lengths):
        with AF.forward_scope(self, x.device):
            return self._forward(x, lengths)

    def _forward(self, x, lengths):
        conv, bn = self.subsampling_module.layers[0][0], self.subsampling_module.layers[0][1]
        if self.training:
            AF.count_batch(bn)
        lengths = torch.div(lengths, 160, rounding_mode="floor") + 1      # preprocessing.py:76-77
        spec = self.spec_augment.params() if (self.training and self.spec_augment.enabled) else None
        mel_len = lengths.to(device=x.device, dtype=torch.long) if spec is not None else None
        x = AF.AudioStemFn.apply(x, self.audio_preprocessing.MelScale.fb, conv.weight, conv.bias, bn.weight, bn.bias,
                                 bn.running_mean, bn.running_var, self.training, bn.momentum, spec, mel_len)
        lengths = torch.div(lengths - 1, 2, rounding_mode="floor") + 1    # modules.py:127
        x = AF.LinearFn.apply(x, self.linear.weight, self.linear.bias, False, self._proj_layout)
        x, lengths, interctc_outputs = self.back_end(x, lengths)
        if not isinstance(self.head, nn.Identity):
            x = AF.LinearFn.apply(x, self.head.weight, self.head.bias, True, None)
        return x, lengths, interctc_outputs


class ResNet(nn.Module):
    """ResNet-18 trunk without stem on channels-last images (networks.py:32-146 with include_stem=False)."""

    def __init__(self, dim_input=3, dim_output=1000, model="ResNet18", include_stem=False, include_head=True):
        super().__init__()
        assert model == "ResNet18" and not include_stem and include_head
        dim_stem, dim_blocks, num_blocks = 64, [64, 128, 256, 512], [2, 2, 2, 2]
        self.stem = nn.Identity()
        self.blocks = nn.ModuleList()
        for stage_id in range(4):
            for block_id in range(num_blocks[stage_id]):
                if block_id == 0:
                    stride = (1, 1) if stage_id == 0 else (2, 2)
                    in_features = dim_stem if stage_id == 0 else dim_blocks[stage_id - 1]
                else:
                    stride, in_features = (1, 1), dim_blocks[stage_id]
                self.blocks.append(ResNetBlock(in_features, dim_blocks[stage_id], (3, 3), stride, "ReLU", True))
        self.head = nn.Sequential(Placeholder("GlobalAvgPool2d"),
                                  Linear(dim_blocks[-1], dim_output, weight_init="he_normal", bias_init="zeros"))

    def forward(self, x):
        x = AF.JoinSideFn.apply(x)       # its backward joins the weight-gradient side stream of the blocks below
        with AF.trunk_scope():
            for block in self.blocks:
                x = block(x)
        x = AF.AvgPoolFn.apply(x)
        return AF.LinearFn.apply(x, self.head[1].weight, self.head[1].bias, False, None)


class _VideoStemParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([nn.Sequential(
            Conv3d(1, 64, (5, 7, 7), stride=(1, 2, 2)), nn.BatchNorm3d(64), Placeholder("ReLU"), Placeholder("Identity"))])


class VisualEfficientConformerEncoder(nn.Module):
    def __init__(self, include_head=True, vocab_size=256, interctc_blocks=[3, 6, 9], num_blocks=[6, 6], loss_prefix="ctc"):
        super().__init__()
        dim_model = [256, 360]
        self.front_end = nn.Sequential(
            _VideoStemParams(),
            Placeholder("MaxPool3d((1,3,3), stride (1,2,2), same) - fused into the stem kernel chain"),
            Placeholder("VideoToImages"),
            ResNet(include_stem=False, dim_output=dim_model[0], model="ResNet18"),
        )
        self.expand_time = Placeholder("ImagesToVideos")
        self.back_end = ConformerInterCTC(
            dim_model=dim_model, num_blocks=num_blocks, interctc_blocks=interctc_blocks, vocab_size=vocab_size,
            att_params=_att("RelPos1dMultiHeadAttention", 4),
            conv_params={"class": "Conv1d", "params": {"padding": "same", "kernel_size": 15}},
            ff_ratio=4, drop_rate=0.1, conv_stride=2, batch_norm=True, loss_prefix=loss_prefix)
        self.head = Linear(dim_model[-1], vocab_size) if include_head else nn.Identity()

    def forward(self, x, lengths):
        """x: (B,1,T,H,W) as the reference passes it (models_zoo.py:109) or (B,T,H,W,1) - identical memory for C = 1."""
        with AF.forward_scope(self, x.device):
            return self._forward(x, lengths)

    def _forward(self, x, lengths):
        if x.dim() == 5 and x.shape[1] == 1 and x.shape[-1] != 1:
            x = x.permute(0, 2, 3, 4, 1)
        B, T = x.shape[0], x.shape[1]
        conv, bn = self.front_end[0].layers[0][0], self.front_end[0].layers[0][1]
        if self.training:
            AF.count_batch(bn)
        x = AF.VideoStemFn.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                 self.training, bn.momentum)
        x = self.front_end[3](x)                 # (B*T, 256)
        x = x.view(B, T, -1)
        x, lengths, interctc_outputs = self.back_end(x, lengths)
        if not isinstance(self.head, nn.Identity):
            x = AF.LinearFn.apply(x, self.head.weight, self.head.bias, True, None)
        return x, lengths, interctc_outputs


class AudioVisualEfficientConformerEncoder(nn.Module):
    def __init__(self, include_head=True, vocab_size=256, v_interctc_blocks=[3, 6], a_interctc_blocks=[8, 11], f_interctc_blocks=[2]):
        super().__init__()
        dim_model = 360
        self.video_encoder = VisualEfficientConformerEncoder(include_head=False, vocab_size=vocab_size,
                                                             interctc_blocks=v_interctc_blocks, num_blocks=[6, 1], loss_prefix="v_ctc")
        self.audio_encoder = AudioEfficientConformerEncoder(include_head=False, vocab_size=vocab_size,
                                                            interctc_blocks=a_interctc_blocks, num_blocks=[5, 6, 1], loss_prefix="a_ctc")
        self.fusion_module = FusionModule(a_dim_model=dim_model, v_dim_model=dim_model, f_dim_model=dim_model)
        self.audio_visual_encoder = ConformerInterCTC(
            dim_model=dim_model, num_blocks=5, interctc_blocks=f_interctc_blocks, vocab_size=vocab_size,
            att_params=_att("RelPos1dMultiHeadAttention", 4),
            conv_params={"class": "Conv1d", "params": {"padding": "same", "kernel_size": 15}},
            ff_ratio=4, drop_rate=0.1, conv_stride=2, batch_norm=True, loss_prefix="f_ctc")
        self.head = Linear(dim_model, vocab_size) if include_head else nn.Identity()

    # The two encoders are independent until the fusion module: the audio branch (~600 small, latency-bound launches) is issued
    # on a second CUDA stream so that it fills the SMs the video front-end's kernels leave idle.  Autograd runs every backward
    # node on its forward's stream, so the two backward passes overlap the same way; under CUDA-graph capture the fork / join
    # becomes two parallel branches of the graph.  Set to False for a single-stream schedule.
    overlap_branches = True
    _side_streams = {}

    def _side_stream(self, device, which=0):
        """which = 0: audio branch, 1: video branch.  Neither launches its GEMMs as programmatic dependents: a grid scheduled early
        parks one CTA per SM (with its shared memory and TMEM) until the kernel in front of it has drained, which is free on a stream
        that owns the GPU but takes those SMs from the other branch while two run side by side (measured: 45.5 -> 46.1 ms).  The
        caller's stream - fusion blocks, heads, losses, and all of the single-branch models - keeps them."""
        s = self._side_streams.get((device, which))
        if s is None:
            s = torch.cuda.Stream(device=device)
            self._side_streams[(device, which)] = s
            if os.environ.get("AVEC_PDL_SIDE", "0") != "1":
                ops.pdl_exclude_stream(s)
        return s

    def forward(self, video, video_len, audio, audio_len):
        with AF.forward_scope(self, video.device):
            return self._forward(video, video_len, audio, audio_len)

    def _forward(self, video, video_len, audio, audio_len):
        if self.overlap_branches and video.is_cuda:
            cur = torch.cuda.current_stream(video.device)
            side_a, side_v = self._side_stream(video.device, 0), self._side_stream(video.device, 1)
            side_a.wait_stream(cur)
            side_v.wait_stream(cur)
            with torch.cuda.stream(side_a):
                audio, audio_len, audio_interctc_outputs = self.audio_encoder(audio, audio_len)
            with torch.cuda.stream(side_v), AF.no_wgrad_overlap():      # the audio branch already fills the gaps of the video branch
                video, video_len, video_interctc_outputs = self.video_encoder(video, video_len)
            cur.wait_stream(side_a)
            cur.wait_stream(side_v)
            # produced on the side streams, consumed (fusion, losses) on the caller's stream
            for t in [audio, video] + [v[0] for v in audio_interctc_outputs.values()] + [v[0] for v in video_interctc_outputs.values()]:
                t.record_stream(cur)
        else:
            video, video_len, video_interctc_outputs = self.video_encoder(video, video_len)
            audio, audio_len, audio_interctc_outputs = self.audio_encoder(audio, audio_len)
        x = self.fusion_module(audio, video)
        lengths = audio_len
        x, lengths, interctc_outputs = self.audio_visual_encoder(x, lengths)
        interctc_outputs.update(video_interctc_outputs)
        interctc_outputs.update(audio_interctc_outputs)
        if not isinstance(self.head, nn.Identity):
            x = AF.LinearFn.apply(x, self.head.weight, self.head.bias, True, None)
        return x, lengths, interctc_outputs
