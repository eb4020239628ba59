"""Generator contract of the scheduler seam (reference sstar/generators/generic_generator.py:23-48):
``get()`` yields one kwargs dict per task, ``__len__`` is the task count."""
from abc import ABC, abstractmethod


class GenericGenerator(ABC):
    @abstractmethod
    def get(self, **kwargs):
        """Yield the keyword arguments of each task."""

    @abstractmethod
    def __len__(self):
        """Number of tasks ``get()`` yields."""
