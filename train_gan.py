"""python train_gan.py --phase pre-train|train-gan|fine-tune -- same entry point as the reference's train_gan.py."""
import argparse

from pnp_b200.train_gan import main

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--phase", type=str, default=None)
    a, rest = p.parse_known_args()
    main(phase=a.phase, argv=["--phase", str(a.phase)] + rest)
