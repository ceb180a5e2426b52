"""2-LSTM FedShakespeare task model (BASELINE config #4)."""
from msrflute_b200.models.rnn_shakespeare import RNN, CharLSTM  # noqa: F401
