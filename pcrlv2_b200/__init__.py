"""pcrlv2_b200 -- B200-native drop-in for the PCRLv2 3-D self-supervised pre-training hot path.

Public surface mirrors the reference: ``pcrlv2_b200.models.PCRLv23d``,
``pcrlv2_b200.train_3d.{train_pcrlv2_3d, train_pcrlv2_inner, cos_loss}``, ``pcrlv2_b200.main``
(CLI) and ``pcrlv2_b200.utils.{adjust_learning_rate, AverageMeter}``.
"""
__version__ = "0.1.0"
