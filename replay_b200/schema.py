"""Minimal stand-ins for ``replay.data.nn.TensorSchema`` / ``TensorFeatureInfo`` exposing only the duck-typed surface
the sequential models touch (SURVEY.md §8b "Duck-type surface"): when RePlay itself is importable, pass its own schema
objects instead - the models accept either."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class TensorFeatureInfo:
    name: str
    cardinality: int
    padding_value: int
    embedding_dim: int
    is_seq: bool = True
    is_cat: bool = True

    def _set_cardinality(self, n: int) -> None:
        self.cardinality = n


class _Single:
    def __init__(self, f):
        self._f = f

    def item(self):
        return self._f


class TensorSchema:
    """One categorical sequential item-id feature (what SASRec / BERT4Rec need on the hot path)."""

    def __init__(self, item_feature: TensorFeatureInfo, query_id_feature_name: str = "query_id",
                 timestamp_feature_name: str | None = None):
        self._item = item_feature
        self.query_id_feature_name = query_id_feature_name
        self.timestamp_feature_name = timestamp_feature_name

    @property
    def item_id_features(self):
        return _Single(self._item)

    @property
    def item_id_feature_name(self) -> str:
        return self._item.name

    def items(self):
        return [(self._item.name, self._item)]

    def __getitem__(self, name):
        if name != self._item.name:
            raise KeyError(name)
        return self._item

    @property
    def categorical_features(self):
        return {self._item.name: self._item}

    @property
    def numerical_features(self):
        return {}


def item_feature_of(schema):
    """(name, cardinality, padding_value, embedding_dim) from a RePlay TensorSchema or the stand-in above."""
    f = schema.item_id_features.item()
    return schema.item_id_feature_name, int(f.cardinality), int(f.padding_value), getattr(f, "embedding_dim", None)
