"""Names of the 11 items DetectionLayer.debug_outputs() returns, in the reference's order (detection.py:268-279)."""
DET_DEBUG_KEYS = ("class_ids", "indices", "mesh", "ixs", "class_scores", "bbox_delta", "refined_proposals",
                  "clipped_proposals_list", "pre_nms_class_ids_list", "pre_nms_scores_list", "pre_nms_proposals_list")
