"""Pre/post-processing helpers (reference: neuralmonkey/processors/helpers.py)."""
from typing import Any, Callable, Dict, Generator, List


def preprocess_char_based(sentence: List[str]) -> List[str]:
    return list(" ".join(sentence))


def preprocess_add_noise(sentence: List[str]) -> List[str]:
    return sentence


def postprocess_char_based(sentences: List[List[str]]) -> List[List[str]]:
    result = []
    for sentence in sentences:
        result.append("".join(sentence).split(" "))
    return result


def untruecase(sentences: List[List[str]]) -> Generator[List[str], None, None]:
    for sentence in sentences:
        if sentence:
            yield [sentence[0].capitalize()] + sentence[1:]
        else:
            yield []


def pipeline(processors: List[Callable]) -> Callable:
    def process(data: Any) -> Any:
        for processor in processors:
            data = processor(data)
        return data
    return process
