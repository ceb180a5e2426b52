from msrflute_b200.data.federated import ArrayDataLoader
from experiments.cv_cnn_femnist.dataloaders.dataset import Dataset


class DataLoader(ArrayDataLoader):
    dataset_class = Dataset
