from .metrics import ItemTable, eval_arrays, eval_model, get_item_embeddings, metrics_topK, print_metrics
from .dataset import BuildTrainDataset
