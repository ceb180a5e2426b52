from msrflute_b200.data.federated import ArrayDataLoader
from experiments.cv_lr_mnist.dataloaders.dataset import Dataset


class DataLoader(ArrayDataLoader):
    dataset_class = Dataset
