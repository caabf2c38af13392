"""Module-level stand-in for hardnetNAS/general_functions/Losses.py, which the supernet imports as a module
(`import general_functions.Losses as Losses`, model_supernet.py:7) and calls as `Losses.loss_HardNet(anchor, positive,
margin)`: same three names and signatures; the loss is the always-swap form on the fused B200 kernel."""
from ..losses import distance_matrix_vector, distance_vectors_pairwise, loss_HardNet_nas


def loss_HardNet(anchor, positive, margin=1.0):
    """hardnetNAS/general_functions/Losses.py:27-51: hardest-in-batch triplet margin loss with anchor swap always on."""
    return loss_HardNet_nas(anchor, positive, margin=margin)


__all__ = ["distance_matrix_vector", "distance_vectors_pairwise", "loss_HardNet"]
