"""ModelPart base classes (reference: neuralmonkey/model/model_part.py:9-103)."""
from typing import Iterable, List, Set, Tuple

from neuralmonkey_b200.model.feedable import Feedable
from neuralmonkey_b200.model.parameterized import InitializerSpecs, Parameterized


class GenericModelPart:
    @property
    def dependencies(self) -> List[str]:
        """Attribute names regarded as dependents (model_part.py:31-35)."""
        return ["encoder", "parent_decoder", "input_sequence", "attentions", "encoders"]

    def get_dependencies(self) -> Tuple[Set[Feedable], Set[Parameterized]]:
        feedables = set()  # type: Set[Feedable]
        parameterizeds = set()  # type: Set[Parameterized]
        if isinstance(self, Feedable):
            feedables.add(self)
        if isinstance(self, Parameterized):
            parameterizeds.add(self)
        for attr in self.dependencies:
            val = getattr(self, attr, None)
            if val is None:
                continue
            if isinstance(val, GenericModelPart):
                deps = [val]
            elif isinstance(val, Iterable):
                deps = [a for a in val if isinstance(a, GenericModelPart)]
            else:
                deps = []
            for dep in deps:
                feeds, params = dep.get_dependencies()
                feedables |= feeds
                parameterizeds |= params
        return feedables, parameterizeds


class ModelPart(Parameterized, GenericModelPart, Feedable):
    def __init__(self, name: str, reuse: "ModelPart" = None, save_checkpoint: str = None,
                 load_checkpoint: str = None, initializers: InitializerSpecs = None) -> None:
        Parameterized.__init__(self, name, reuse, save_checkpoint, load_checkpoint, initializers)
        GenericModelPart.__init__(self)
        Feedable.__init__(self)
