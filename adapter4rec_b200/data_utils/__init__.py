from .metrics import ItemTable, eval_arrays, eval_model, get_item_embeddings, metrics_topK, print_metrics
from .dataset import BuildTrainDataset
from .preprocess import get_doc_input_bert, load_text_data, read_behaviors, read_images, read_news, read_news_bert
