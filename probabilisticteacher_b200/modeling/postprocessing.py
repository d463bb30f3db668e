"""`detector_postprocess` of the eval-mode path (`pt/modeling/meta_arch/rcnn.py:33-34` ->
detectron2 v0.5 `GeneralizedRCNN.inference(do_postprocess=True)` -> `_postprocess`): detections computed at the
network's input resolution are rescaled to the resolution the caller asked for (`"height"` / `"width"` of the input
dict, the ORIGINAL image size written by the dataset mapper, `pt/data/dataset_mapper.py:162-169`), clipped, and empty
boxes are dropped. Box-only restatement (MASK_ON / KEYPOINT_ON are False on this path). Host-side glue on the
<= 100 detections per image of the pseudo-label filter; nothing here is on the training step."""
import torch

from ..structures import Boxes, FreeInstances


def detector_postprocess(results, output_height, output_width):
    """results: (Free)Instances with `pred_boxes` or `proposal_boxes` at `results.image_size`; a fixed-capacity
    instance (device-side count) is trimmed first (one host sync). Returns a NEW exact-length instance with
    image_size (output_height, output_width); the input's tensors are not modified."""
    if isinstance(results, FreeInstances):
        results = results.trim()
    if isinstance(output_width, torch.Tensor):
        output_width, output_height = float(output_width), float(output_height)
    h, w = results.image_size
    scale_x, scale_y = output_width / w, output_height / h
    out = FreeInstances((int(output_height), int(output_width)))
    for k, v in results.get_fields().items():
        out.set(k, v)
    if out.has("pred_boxes"):
        key = "pred_boxes"
    elif out.has("proposal_boxes"):
        key = "proposal_boxes"
    else:
        raise AssertionError("Predictions must contain boxes!")
    b = out.get(key).tensor.clone()
    b[:, 0::2] *= scale_x
    b[:, 1::2] *= scale_y
    boxes = Boxes(b)
    boxes.clip(out.image_size)
    out.set(key, boxes)
    return out[boxes.nonempty()]


def postprocess_batch(instances, batched_inputs, image_sizes):
    """d2 `GeneralizedRCNN._postprocess`: one `{"instances": ...}` dict per image, at the size named by the input
    dict (default: the network input size)."""
    processed = []
    for res, inp, size in zip(instances, batched_inputs, image_sizes):
        height = inp.get("height", size[0])
        width = inp.get("width", size[1])
        processed.append({"instances": detector_postprocess(res, height, width)})
    return processed
