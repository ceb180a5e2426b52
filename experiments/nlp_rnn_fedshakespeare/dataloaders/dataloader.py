from msrflute_b200.data.federated import ArrayDataLoader
from experiments.nlp_rnn_fedshakespeare.dataloaders.dataset import Dataset


class DataLoader(ArrayDataLoader):
    dataset_class = Dataset
