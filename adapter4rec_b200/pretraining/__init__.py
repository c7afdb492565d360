"""Source-domain stage of the reference (Pretraining/Text, Pretraining/CV): the same TransRec model trained WITHOUT adapters —
the tail of the modality encoder, its projection and the user encoder are fine-tuned; the checkpoint it writes is what the
downstream scripts load through --pretrained_model_name / --pretrained_recsys_model before inserting adapters."""
