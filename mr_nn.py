"""`python mr_nn.py --tables 2 4 [-v]` -- same entry point as the reference's mr_nn.py:121-168."""
import sys

from mr_gan_b200.mr_nn import main, mr_nn  # noqa: F401

if __name__ == '__main__':
    sys.exit(main())
