"""Series pre/post-processors named by the reference's INIs
(behaviour of neuralmonkey/processors/helpers.py).

    preprocess_char_based   words -> characters, with a space token between words
    postprocess_char_based  inverse, for batches of decoded sentences
    untruecase              capitalise the first token of every sentence (lazy)
    pipeline                left-to-right composition of processors
"""
from functools import reduce
from typing import Any, Callable, Iterator, List, Sequence

Sentence = List[str]


def preprocess_char_based(sentence: Sentence) -> Sentence:
    chars = []  # type: Sentence
    for position, word in enumerate(sentence):
        if position:
            chars.append(" ")
        chars.extend(word)
    return chars


def preprocess_add_noise(sentence: Sentence) -> Sentence:
    """Placeholder kept for configuration compatibility: the identity."""
    return sentence


def postprocess_char_based(sentences: Sequence[Sentence]) -> List[Sentence]:
    return ["".join(characters).split(" ") for characters in sentences]


def untruecase(sentences: Sequence[Sentence]) -> Iterator[Sentence]:
    return ([tokens[0].capitalize(), *tokens[1:]] if tokens else [] for tokens in sentences)


def pipeline(processors: Sequence[Callable[[Any], Any]]) -> Callable[[Any], Any]:
    return lambda data: reduce(lambda value, stage: stage(value), processors, data)
