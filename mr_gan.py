"""`python mr_gan.py --tables 1 3 5 6 [-v]` -- same entry point as the reference's mr_gan.py:236-341."""
import sys

from mr_gan_b200.mr_gan import dataset, main, mr_gan  # noqa: F401

if __name__ == '__main__':
    sys.exit(main())
