"""CPU twin of nesie_b200.votenet.VoteNetHarness: identical torch glue, hot-path hooks routed to
the oracle (C restatement of the reference kernels + torch-CPU MLPs).  TEST INFRASTRUCTURE ONLY:
used by parity tests and by bench.py's cpu_baseline / `--impl reference` legs."""
from nesie_b200.votenet import VoteNetHarness

from . import modules as om
from . import restate


class VoteNetOracle(VoteNetHarness):

    @staticmethod
    def _side_pooling_cls():
        from .side_pooling_ref import SidePoolingOracle
        return SidePoolingOracle

    def _backbone(self, points, fps_indices=None, after_level=None):
        return om.backbone_forward(self.backbone, points)

    def _aggregate(self, xyz, feats):
        return om.sa_forward(self.vote_aggregation, xyz, feats)

    def _side_loss(self, surface_pred, box_targets, side_scores, sem_scores, weight):
        return restate.side_uncertainty_loss(surface_pred, box_targets, side_scores, sem_scores,
                                             weight, 10.0, self.alpha)
