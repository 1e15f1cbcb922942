from .proposal_utils import find_top_rpn_proposals, find_top_rpn_proposals_device

__all__ = ["find_top_rpn_proposals", "find_top_rpn_proposals_device"]
