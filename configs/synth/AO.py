"""synthetic twin of configs/LRS23/AO/EffConfInterCTC.py (reference) on the avec_b200 encoders"""
from configs.synth.common import *  # noqa: F401,F403
from configs.synth.common import SyntheticAV, callback_root, nnet, os

vocab_size = 256
callback_path = os.path.join(callback_root, "AO")

model = nnet.AudioEfficientConformerInterCTC(vocab_size=vocab_size, att_type="patch", interctc_blocks=[3, 6, 10, 13])
model.compile(losses=nnet.CTCLoss(zero_infinity=True, assert_shorter=False), loss_weights=[0.5 / 4, 0.5 / 4, 0.5 / 4, 0.5 / 4, 0.5])

collate_fn = nnet.CollateFn(inputs_params=[{"axis": 1, "padding": True}, {"axis": 4}], targets_params=({"axis": 2, "padding": True}, {"axis": 5}))
training_dataset = SyntheticAV(batch_size=4, collate_fn=collate_fn, n=16, seed=1)
evaluation_dataset = SyntheticAV(batch_size=4, collate_fn=collate_fn, n=8, seed=2)
