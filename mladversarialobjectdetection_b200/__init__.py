"""B200-native EOT patch-attack hot path (drop-in for the reference's Patcher / Masker /
BrightnessMatcher layers and PatchAttacker.train_step)."""
__version__ = "0.1.0"
