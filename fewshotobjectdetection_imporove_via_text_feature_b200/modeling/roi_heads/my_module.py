"""Knowledge-distillation losses used by the teacher/student heads.

Mirror of the two functions of defrcn/modeling/roi_heads/my_module.py that the hot path references
(`loss_fn_kd` :393-406, `loss_fn_kd_only` :409-437, called from roi_heads.py:760); the remaining ~1300 lines of that
file (optimal-transport layer, memory banks, generators) are unused experiment code and out of scope.
"""
import torch
import torch.nn.functional as F


def loss_fn_kd(outputs, labels, teacher_outputs, params):
    """alpha*T^2*KL(student/T || teacher/T) (elementwise-mean reduction, as nn.KLDivLoss() defaults) + (1-alpha)*CE."""
    alpha, T = params["alpha"], params["temperature"]
    kl = F.kl_div(F.log_softmax(outputs / T, dim=1), F.softmax(teacher_outputs / T, dim=1), reduction="none").mean()
    return kl * (alpha * T * T) + F.cross_entropy(outputs, labels) * (1.0 - alpha)


def loss_fn_kd_only(outputs, labels, bg_label, teacher_outputs, params):
    """Per-row KL summed over classes, background rows weighted 1.5x, mean over rows, times T^2*alpha."""
    alpha, T = params["alpha"], params["temperature"]
    kl = F.kl_div(F.log_softmax(outputs / T, dim=1), F.softmax(teacher_outputs / T, dim=1), reduction="none").sum(1)
    kl = torch.where(labels == bg_label, kl * 1.5, kl)
    return kl.sum() / labels.shape[0] * T * T * alpha
