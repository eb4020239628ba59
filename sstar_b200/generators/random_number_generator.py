"""Replicate / seed generator of the simulate path (reference
sstar/generators/random_number_generator.py:24-84): ``np.random.seed(seed)`` then
``randint(1, 2**31, end_rep)`` sliced to ``[start_rep, end_rep)``."""
import numpy as np

from .generic_generator import GenericGenerator


class RandomNumberGenerator(GenericGenerator):
    def __init__(self, nrep: int, start_rep: int = 0, seed: int = None):
        if nrep <= 0:
            raise ValueError("nrep must be greater than 0.")
        if start_rep < 0:
            raise ValueError("start_rep must not be negative.")
        end_rep = start_rep + nrep
        self.rep_list = list(range(start_rep, end_rep))
        if seed is None:
            self.seed_list = [None] * nrep
        elif end_rep > 1:
            np.random.seed(seed)
            self.seed_list = np.random.randint(1, 2**31, end_rep).tolist()[start_rep:end_rep]
        else:
            self.seed_list = [seed]

    def get(self):
        for rep, seed in zip(self.rep_list, self.seed_list):
            yield {"rep": rep, "seed": seed}

    def __len__(self):
        return len(self.rep_list)
