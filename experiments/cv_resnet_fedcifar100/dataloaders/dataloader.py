from msrflute_b200.data.federated import ArrayDataLoader
from experiments.cv_resnet_fedcifar100.dataloaders.dataset import Dataset


class DataLoader(ArrayDataLoader):
    dataset_class = Dataset
