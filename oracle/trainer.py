"""Plain-PyTorch CPU restatement of the loss / regulariser / step of
``learning/runModel.py`` (TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def cell_loss(logits, gt, weight_col, loss="kl", cell_norm=None, cell_type="vol"):
    """``Trainer.calcLossAndOA`` cell branch, ``runModel.py:163-211``.

    ``gt`` = ``data.batch_gt`` (``[:, :2]`` = inside/outside percentages), ``weight_col`` =
    ``data.batch_x[:, 0]`` (raw volume).  Returns ``(loss, sum_weighted_loss, sum_weight)``.
    """
    if loss == "kl":  # :171-175
        cl = F.kl_div(F.log_softmax(logits, dim=-1), gt[:, :2], reduction='none')
        cl = torch.sum(cl, dim=1)
    elif loss == "bce":  # :181-183
        cl = F.binary_cross_entropy_with_logits(logits.squeeze(dim=-1), gt[:, 3], reduction='none')
    elif loss == "mse":  # :187-188  (a scalar: F.mse_loss default reduction is 'mean')
        cl = F.mse_loss(torch.sigmoid(logits).squeeze(), gt[:, 0])
    else:
        raise ValueError(loss)
    if cell_norm == "log":  # :193-202
        w = torch.log(1 + weight_col)
    elif cell_norm == "sqrt":
        w = torch.sqrt(weight_col)
    elif cell_type:
        w = weight_col
    else:
        w = torch.ones(size=cl.shape)
    cl = cl * w  # :205
    return cl.sum() / w.sum(), cl.sum(), w.sum()  # :208-211


def overall_accuracy_count(logits, gt):
    """``runModel.py:178-180`` (note: compares inside>outside with argmax==1, i.e. the
    complement of the accuracy — reproduced as is)."""
    return int(torch.sum((gt[:, 0] > gt[:, 1]).type(torch.int64) ==
                         F.log_softmax(logits, dim=-1).argmax(1)).item())


def edge_regularization(logits, edge_index, edge_weight):
    """``Trainer.calcRegularization``, ``runModel.py:109-160`` (unbatched branch
    ``:125-130``): mean over edges of ``|p0[src] - p0[tgt]| * edge_weight``."""
    p = F.softmax(logits, dim=-1)
    d = torch.abs(p[edge_index[0, :]][:, 0] - p[edge_index[1, :]][:, 0])
    return (d * edge_weight).mean()


def train_step(model, optimizer, data, clf):
    """``Trainer.train``, ``runModel.py:264-282`` for the Static model."""
    model.train()
    logits = model(data)
    n_sup = data.batch_adjs[model.num_layers - 1][2][1]
    batch_x = data.all.x[data.batch_n_id[:n_sup]]
    batch_gt = data.all.y[data.batch_n_id[:n_sup]]
    loss, _, _ = cell_loss(logits, batch_gt, batch_x[:, 0], clf.training.loss,
                           clf.regularization.cell_norm, clf.regularization.cell_type)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach(), logits.detach()


def labels(logits):
    """``processing/generate_mesh.py:75``."""
    return F.log_softmax(logits, dim=-1).argmax(1)
