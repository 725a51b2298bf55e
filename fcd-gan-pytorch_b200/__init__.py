"""fcd-gan-pytorch_b200 — B200-native hot path of FCD-GAN (networks + loss stack) behind the reference's
nn.Module call surface.  Import as `fcdgan_b200`."""
__version__ = "0.1.0"
