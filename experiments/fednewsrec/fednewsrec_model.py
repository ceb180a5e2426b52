from msrflute_b200.models.newsrec import (AttentivePooling, Attention, DocEncoder, UserEncoder, FedNewsRec,  # noqa: F401
                                          npratio)
