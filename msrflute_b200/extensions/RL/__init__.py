from .RL import RL, NeuralNetwork, BatchRNN, SequenceWise  # noqa: F401
