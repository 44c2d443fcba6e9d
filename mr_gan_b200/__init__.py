"""mr-gan-b200: B200-native training step of Healthcare-Robotics/mr-gan behind the
reference's own ``mr_gan()`` / ``mr_nn()`` / ``--tables`` surface.

    from mr_gan_b200.mr_gan import dataset, mr_gan      # as `from mr_gan import ...` in the reference
    from mr_gan_b200.mr_nn import mr_nn
"""
from .engine import FoldGroup, MrganError  # noqa: F401
