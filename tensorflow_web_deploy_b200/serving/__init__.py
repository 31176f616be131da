"""Serving front-end (BASELINE configs[4]): a py3 / werkzeug port of the reference's three HTTP routes
(/root/reference/PyClassification.py:29-113), pre-fork workers, and the cross-process micro-batcher that feeds one
GPU-owner process per GPU."""
