"""Command-line flags of the reference (utility/parser.py:4-56), same names, types and defaults.

Table-driven restatement: (flag, kwargs) rows instead of one add_argument call per line.
`--mask` keeps the reference's `type=bool` quirk (any non-empty string is True, parser.py:39).
`--dataset netflix|movielens` are accepted as aliases of the on-disk directory names
(README.md:80-82 documents the short names; main.py:69-72 only handles the long ones).
"""
import argparse

_FLAGS = [
    ("data_path", dict(nargs="?", default="./data/", help="Input data path")),
    ("seed", dict(type=int, default=2022, help="Random seed")),
    ("dataset", dict(nargs="?", default="netflix", help="Choose a dataset from {movieLens, netflix}")),
    ("verbose", dict(type=int, default=5, help="Interval of evaluation.")),
    ("epoch", dict(type=int, default=1000, help="Number of epoch.")),
    ("regs", dict(nargs="?", default="[1e-5,1e-5,1e-2]", help="Regularizations.")),
    ("embed_size", dict(type=int, default=64, help="Embedding size.")),
    ("weight_size", dict(nargs="?", default="[64, 64]", help="Output sizes of every layer")),
    ("early_stopping_patience", dict(type=int, default=7, help="Early Stop Patience")),
    ("mess_dropout", dict(nargs="?", default="[0.1, 0.1]", help="Keep probability w.r.t. message dropout")),
    ("sparse", dict(type=int, default=1, help="Sparse or dense adjacency matrix")),
    ("debug", dict(action="store_true")),
    ("norm_type", dict(nargs="?", default="sym", help="Adjacency matrix normalization operation")),
    ("gpu_id", dict(type=int, default=0, help="GPU ID")),
    ("Ks", dict(nargs="?", default="[10, 20, 50]", help="K value of ndcg/recall @ k")),
    ("test_flag", dict(nargs="?", default="part", help="Specify the test type from {part, full}")),
    ("sc", dict(type=float, default=1.0, help="GCN self connection")),
    ("feat_reg_decay", dict(default=1e-5, type=float, help="Feature Reg Decay")),
    ("title", dict(default="try_to_draw_line", type=str, help="")),
    ("cf_model", dict(nargs="?", default="lightgcn", help="Downstream Collaborative Filtering model")),
    ("point", dict(default="", type=str, help="")),
    # train
    ("batch_size", dict(type=int, default=1024, help="Batch size.")),
    ("lr", dict(type=float, default=0.0001, help="Learning rate.")),
    ("de_lr", dict(type=float, default=0.0002, help="Decoder learning rate.")),
    ("weight_decay", dict(default=1e-4, type=float, help="Weight_decay")),
    # model
    ("layers", dict(type=int, default=1, help="Number of graph conv layers")),
    ("drop_rate", dict(type=float, default=0.0, help="Dropout rate")),
    ("mask_rate", dict(type=float, default=0.0, help="Mask rate")),
    ("mask", dict(type=bool, default=False, help="If mask")),
    ("user_cat_rate", dict(type=float, default=2.8, help="User cat rate")),
    ("item_cat_rate", dict(type=float, default=0.005, help="Item cat rate")),
    ("model_cat_rate", dict(type=float, default=0.02, help="Model cat rate")),
    ("de_drop1", dict(default=0.31, type=float, help="for D model2")),
    ("de_drop2", dict(default=0.5, type=float, help="")),
    # loss
    ("aug_mf_rate", dict(type=float, default=0.012, help="Augmentation mf rate")),
    ("prune_loss_drop_rate", dict(type=float, default=0.71, help="Prune loss drop rate")),
    ("mm_mf_rate", dict(type=float, default=0.0001, help="MM mf rate")),
    ("feat_loss_type", dict(default="sce", type=str, help="Feature loss type")),
    ("att_re_rate", dict(type=float, default=0.00000, help="Attribute restoration rate")),
    ("alpha_l", dict(type=float, default=2, help="`pow`inddex for `sce` loss")),
    ("aug_sample_rate", dict(type=float, default=0.1, help="Augmentation sample rate")),
    ("mf_emb_rate", dict(type=float, default=0.0, help="MF embedding rate")),
]

# B200-side extras (not in the reference; all optional)
_EXTRA = [
    ("proj_mode", dict(default="3xtf32", choices=["3xtf32", "tf32", "fp32"], help="tensor-core mode of the projection / scoring GEMMs")),
    ("feat_layout", dict(default="rows", choices=["rows", "panels"], help="HBM layout of the constant side-feature tables: row-major, or 32-column panels (contiguous tiles for the projection kernels)")),
    ("cuda_graph", dict(type=int, default=1, help="replay the training step from a CUDA graph (1) or launch eagerly (0)")),
    ("host_sampler", dict(default="native", choices=["native", "python"], help="bit-identical C sampler or the reference's Python loops")),
]

DATASET_ALIASES = {"netflix": "netflix_valid_item", "movielens": "preprocessed_raw_MovieLens", "movieLens": "preprocessed_raw_MovieLens"}


def build_parser():
    ap = argparse.ArgumentParser(description="")
    for name, kw in _FLAGS + _EXTRA:
        ap.add_argument("--" + name, **kw)
    return ap


def parse_args(argv=None):
    """parse_args() -> Namespace (reference: utility/parser.py:4).  Unknown flags are an error."""
    return build_parser().parse_args(argv)


def resolve_dataset_dir(data_path, dataset):
    """Directory that holds the dataset: the literal name if it exists, else its alias."""
    import os
    first = os.path.join(data_path, dataset)
    if os.path.isdir(first):
        return first
    alias = DATASET_ALIASES.get(dataset)
    if alias and os.path.isdir(os.path.join(data_path, alias)):
        return os.path.join(data_path, alias)
    return first
