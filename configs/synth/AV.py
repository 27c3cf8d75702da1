"""synthetic twin of configs/LRS23/AV/EffConfInterCTC.py (reference) on the avec_b200 encoders"""
from configs.synth.common import *  # noqa: F401,F403
from configs.synth.common import SyntheticAV, callback_root, nnet, os

vocab_size = 256
loss_weights = {"v_ctc_2": 0.5 / 3, "v_ctc_5": 0.5 / 3, "a_ctc_7": 0.5 / 3, "a_ctc_10": 0.5 / 3, "f_ctc_1": 0.5 / 3, "outputs": 0.5}
callback_path = os.path.join(callback_root, "AV")

model = nnet.AudioVisualEfficientConformerInterCTC(vocab_size=vocab_size, v_interctc_blocks=[3, 6], a_interctc_blocks=[8, 11], f_interctc_blocks=[2])
model.compile(losses=nnet.CTCLoss(zero_infinity=True, assert_shorter=False), loss_weights=loss_weights)

# inputs (video, video_len, audio, audio_len), targets (label, label_len): the axis mapping of the real AV config
collate_fn = nnet.CollateFn(inputs_params=[{"axis": 0, "padding": True}, {"axis": 3}, {"axis": 1, "padding": True}, {"axis": 4}],
                            targets_params=({"axis": 2, "padding": True}, {"axis": 5}))
training_dataset = SyntheticAV(batch_size=4, collate_fn=collate_fn, n=16, seed=1)
evaluation_dataset = SyntheticAV(batch_size=4, collate_fn=collate_fn, n=8, seed=2)
