"""Sample pairing for GAN training (reference: gan/gan_sampling_methods.py).  Only the DummySampler — the reference's
one known-answer fixture (:191-201) — is kept; the scene-scanning samplers are host-side data preparation (SURVEY §2
row 7: out of scope)."""
from abc import ABC, abstractmethod

import numpy


class Sampler(ABC):
    @abstractmethod
    def get_sample_pairs(self, data_set, loader, shadow_map):
        pass


class DummySampler(Sampler):
    def __init__(self, element_count, fill_value, coefficient):
        self._element_count = element_count
        self._fill_value = fill_value
        self._coefficient = coefficient

    def get_sample_pairs(self, data_set, loader, shadow_map):
        data_shape_info = data_set.get_data_shape()
        shadow_data_as_matrix = numpy.full(numpy.concatenate([[self._element_count], data_shape_info]),
                                           fill_value=self._fill_value, dtype=numpy.float32)
        return shadow_data_as_matrix * self._coefficient, shadow_data_as_matrix
