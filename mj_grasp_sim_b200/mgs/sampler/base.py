"""GraspGenerator protocol (/root/reference/mgs/sampler/base.py:22-32)."""
from abc import ABC, abstractmethod


class GraspGenerator(ABC):
    def __init__(self, obj):
        self.obj = obj

    @abstractmethod
    def generate_grasps(self, num):
        """Generate grasps for the object."""
