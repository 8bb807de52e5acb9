"""Abstract layer API of the reference (inf/layers/flowlayer.py:7-51): the contract that
FlowSequential (inf/layers/flowsequential.py:20-43) relies on."""
from abc import ABCMeta, abstractmethod

import torch.nn as nn


class FlowLayer(nn.Module, metaclass=ABCMeta):

    @abstractmethod
    def forward(self, input, context=None):
        """-> (output, log|det J| per sample or 0.0)"""

    @abstractmethod
    def reverse(self, input, context=None):
        """-> input of forward"""

    @abstractmethod
    def logdet(self, input, context=None):
        """-> log|det J|"""


def mark_expensive(func):
    func._expensive_computation = True
    return func
