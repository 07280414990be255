"""Oracle: ODISE / Mask2Former inference post-processing (CPU, torch).  TEST INFRASTRUCTURE ONLY.

Restates the tail of CategoryODISE.forward without a clip_head (odise/modeling/meta_arch/odise.py:326-370) and
MaskFormer.semantic_inference / panoptic_inference / instance_inference
(third_party/Mask2Former/mask2former/maskformer_model.py:280-380).  Pinned: tests/test_oracle_cpu.py calls the
reference's own methods (imported through oracle/refshim.py) on the same inputs.
"""
import torch
import torch.nn.functional as F


def upsample_masks(mask_pred, size):
    """odise.py:326-331: F.interpolate(pred_masks, size=(H, W), mode="bilinear", align_corners=False)."""
    return F.interpolate(mask_pred, size=size, mode="bilinear", align_corners=False)


def sem_seg_postprocess(result, img_size, output_height, output_width):
    """detectron2 v0.6 modeling/postprocessing.py::sem_seg_postprocess (un-vendored; restated, unpinned): crop the padded
    prediction to the image, bilinear resize to the requested size.  result [C, Hpad, Wpad]."""
    result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]


def semantic_inference(mask_cls, mask_pred):
    """maskformer_model.py:280-284.  mask_cls [Q, K+1], mask_pred [Q, H, W] logits -> [K, H, W]."""
    mask_cls = F.softmax(mask_cls, dim=-1)[..., :-1]
    mask_pred = mask_pred.sigmoid()
    return torch.einsum("qc,qhw->chw", mask_cls, mask_pred)


def panoptic_inference(mask_cls, mask_pred, num_classes, thing_ids, object_mask_threshold=0.0, overlap_threshold=0.8):
    """maskformer_model.py:286-342.  Returns (panoptic_seg int32 [H, W], segments_info list of dicts)."""
    scores, labels = F.softmax(mask_cls, dim=-1).max(-1)
    mask_pred = mask_pred.sigmoid()
    keep = labels.ne(num_classes) & (scores > object_mask_threshold)
    cur_scores, cur_classes, cur_masks = scores[keep], labels[keep], mask_pred[keep]
    cur_prob_masks = cur_scores.view(-1, 1, 1) * cur_masks
    h, w = cur_masks.shape[-2:]
    panoptic_seg = torch.zeros((h, w), dtype=torch.int32)
    segments_info = []
    current_segment_id = 0
    if cur_masks.shape[0] == 0:
        return panoptic_seg, segments_info
    cur_mask_ids = cur_prob_masks.argmax(0)
    stuff_memory_list = {}
    thing_ids = set(int(t) for t in thing_ids)
    for k in range(cur_classes.shape[0]):
        pred_class = cur_classes[k].item()
        isthing = pred_class in thing_ids
        mask_area = (cur_mask_ids == k).sum().item()
        original_area = (cur_masks[k] >= 0.5).sum().item()
        mask = (cur_mask_ids == k) & (cur_masks[k] >= 0.5)
        if mask_area > 0 and original_area > 0 and mask.sum().item() > 0:
            if mask_area / original_area < overlap_threshold:
                continue
            if not isthing:
                if int(pred_class) in stuff_memory_list.keys():
                    panoptic_seg[mask] = stuff_memory_list[int(pred_class)]
                    continue
                else:
                    stuff_memory_list[int(pred_class)] = current_segment_id + 1
            current_segment_id += 1
            panoptic_seg[mask] = current_segment_id
            segments_info.append({"id": current_segment_id, "isthing": bool(isthing), "category_id": int(pred_class)})
    return panoptic_seg, segments_info


def instance_inference(mask_cls, mask_pred, num_classes, thing_ids, topk=100, panoptic_on=True):
    """maskformer_model.py:344-380 without the detectron2 Instances container:
    returns dict(pred_masks [N, H, W] float 0/1, scores [N], pred_classes [N])."""
    Q = mask_cls.shape[0]
    scores = F.softmax(mask_cls, dim=-1)[:, :-1]
    labels = torch.arange(num_classes).unsqueeze(0).repeat(Q, 1).flatten(0, 1)
    scores_per_image, topk_indices = scores.flatten(0, 1).topk(topk, sorted=False)
    labels_per_image = labels[topk_indices]
    topk_indices = topk_indices // num_classes
    mp = mask_pred[topk_indices]
    if panoptic_on:
        thing = torch.tensor(sorted(int(t) for t in thing_ids))
        keep = torch.isin(labels_per_image, thing)
        scores_per_image, labels_per_image, mp = scores_per_image[keep], labels_per_image[keep], mp[keep]
    pm = (mp > 0).float()
    mask_scores = (mp.sigmoid().flatten(1) * pm.flatten(1)).sum(1) / (pm.flatten(1).sum(1) + 1e-6)
    return dict(pred_masks=pm, scores=scores_per_image * mask_scores, pred_classes=labels_per_image)
