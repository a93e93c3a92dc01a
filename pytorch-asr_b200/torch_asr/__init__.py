"""Namespace of the native modules, named as the reference names its own
(`torch_asr._latgen_lib`, asr/kaldi/setup.py:57-59): `torch_asr._ctc_lib`."""
