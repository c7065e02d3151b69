"""Drop-in for the reference's TSP/train.py: reads config.yml from the current directory (torchrun for multi-GPU)."""
from ..train_loop import main

if __name__ == "__main__":
    main("tsp")
