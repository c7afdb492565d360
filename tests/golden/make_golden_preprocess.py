"""Runs the UNMODIFIED reference readers (Downstream/Text/data_utils/preprocess.py: read_news_bert, read_behaviors,
get_doc_input_bert) on the fixture of preprocess_fixture.py and stores what they return in tests/golden/preprocess/golden.json.
preprocess.py imports only numpy and torch, so it is executed from its own file (the package __init__ would pull the dataset /
metrics modules in, which is unnecessary here)."""
import importlib.util
import json
import logging
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import preprocess_fixture as F  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_preprocess", "/root/reference/Downstream/Text/data_utils/preprocess.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

F.write_fixture()
args = types.SimpleNamespace(news_attributes=['title'], num_words_title=F.NUM_WORDS, num_words_abstract=50, num_words_body=50)
log = logging.getLogger("golden_preprocess")
before_dic, before_name_to_id = ref.read_news_bert(F.NEWS, args, F.toy_tokenizer)
item_num, item_id_to_dic, tr, va, te, hv, ht = ref.read_behaviors(F.BEHAVIORS, before_dic, before_name_to_id, F.MAX_SEQ_LEN,
                                                                    F.MIN_SEQ_LEN, log)
title, mask, *rest = ref.get_doc_input_bert(item_id_to_dic, args)
assert all(r is None for r in rest)
item_content = np.concatenate([title, mask], axis=1)
out = {"before_item_num": len(before_name_to_id), "item_num": item_num, "item_content_dtype": str(item_content.dtype),
       "item_content": item_content.tolist(), "users_train": tr, "users_valid": va, "users_test": te,
       "users_history_for_valid": {k: v.tolist() for k, v in hv.items()},
       "users_history_for_test": {k: v.tolist() for k, v in ht.items()},
       "read_news": {"ids": len(ref.read_news(F.NEWS)[0]), "first": ref.read_news(F.NEWS)[0][1]}}
with open(os.path.join(F.DIR, "golden.json"), "w") as f:
    json.dump(out, f, sort_keys=True)
    f.write("\n")
print("items", item_num, "of", len(before_name_to_id), "users", len(tr))
