"""Drop-in module path for gauche/kernels/fingerprint_kernels/dice_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import DiceKernel, batch_dice_sim

__all__ = ["DiceKernel", "batch_dice_sim"]
