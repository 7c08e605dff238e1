"""Evaluation model wrapper (reference ``src/metrics/preparation.py:43-122``): ``LoadEvalModel.get_outputs(x, quantize)``
-> (pool features [B,2048], logits [B,1008]) for the InceptionV3_tf backbone with the 'legacy' (torch bilinear) or
'friendly' (PIL bilinear) post-resizer (src/utils/resize.py:50-94).  The
reference quantises on the host, resizes image by image in Python and copies back; here the whole pre-processing is one
device kernel feeding the Inception conv pipeline."""
import torch

from .inception_net import InceptionV3, seeded_state_dict


class LoadEvalModel(object):
    def __init__(self, eval_backbone, post_resizer, world_size, distributed_data_parallel, device, state_dict=None):
        if eval_backbone != "InceptionV3_tf":
            raise NotImplementedError("only the InceptionV3_tf backbone is on the sgb200 hot path (SURVEY.md section 8, a17)")
        if post_resizer not in ("legacy", "friendly"):
            raise NotImplementedError("post_resizer '%s': the device kernel implements 'legacy' (torch bilinear) and 'friendly' "
                                      "(PIL bilinear); 'clean' (PIL bicubic) is not on the hot path" % post_resizer)
        self.eval_backbone, self.post_resizer, self.device = eval_backbone, post_resizer, device
        self.res = 299
        self.pretrained = state_dict is not None
        # the FID weights file cannot be downloaded in this environment: fall back to seeded weights (stated in DESIGN.md)
        self.model = InceptionV3(state_dict if state_dict is not None else seeded_state_dict(0), device)

    def eval(self):
        pass

    def get_outputs(self, x, quantize=False):
        """x: NCHW images; quantize=True: float in [-1,1] (generated images); False: values already in 0..255."""
        return self.model.forward(x.to(self.device, torch.float32), quantize=quantize, resizer=self.post_resizer)
