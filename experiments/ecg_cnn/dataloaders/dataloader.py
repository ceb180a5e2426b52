from msrflute_b200.data.federated import ArrayDataLoader
from experiments.ecg_cnn.dataloaders.dataset import Dataset


class DataLoader(ArrayDataLoader):
    dataset_class = Dataset
