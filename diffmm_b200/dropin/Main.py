"""Accelerated driver with the reference's CLI: ``python Main.py -c conf/tiktok.toml``."""
import _path  # noqa: F401
from diffmm_b200.Main import Coach, main, seed_it  # noqa: F401

if __name__ == "__main__":
    main()
