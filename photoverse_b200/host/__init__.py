"""Host-side callers of the hot path (plain PyTorch): SD-1.5-shaped UNet, DDIM schedule, generation loop."""
