"""Data carriers passed between the four model stages (embed -> encode -> modify -> project).
Same field names as the reference's ``models/common/layers/data/sequence.py:8-112`` so that code
written against the reference (training modules, evaluators) works unchanged."""
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch


@dataclass
class InputSequence:
    sequence: torch.Tensor                 # (N,S) item ids (or (N,S,BS) baskets)
    padding_mask: Optional[torch.Tensor]   # (N,S) True where a real item sits
    attributes: Dict[str, Any]

    def get_attributes(self) -> List[str]:
        return list(self.attributes.keys())

    def has_attribute(self, name: str) -> bool:
        return name in self.attributes

    def get_attribute(self, name: str) -> Optional[Any]:
        return self.attributes.get(name)

    def set_attribute(self, name: str, value: Any, overwrite: bool = False):
        if name in self.attributes and not overwrite:
            raise Exception("Attribute is already set.")
        self.attributes[name] = value


@dataclass
class EmbeddedElementsSequence:
    embedded_sequence: torch.Tensor        # (N,S,H)
    input_sequence: Optional[InputSequence] = None


@dataclass
class SequenceRepresentation:
    encoded_sequence: torch.Tensor         # (N,S,H)
    embedded_elements_sequence: Optional[EmbeddedElementsSequence] = None

    @property
    def input_sequence(self):
        e = self.embedded_elements_sequence
        return None if e is None else e.input_sequence


@dataclass
class ModifiedSequenceRepresentation:
    modified_encoded_sequence: torch.Tensor
    sequence_representation: Optional[SequenceRepresentation] = None

    @property
    def input_sequence(self):
        r = self.sequence_representation
        return None if r is None else r.input_sequence
