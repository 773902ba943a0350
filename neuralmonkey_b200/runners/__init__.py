from neuralmonkey_b200.runners.runner import GreedyRunner
