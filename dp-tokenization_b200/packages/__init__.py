"""Drop-in mirror of the reference's ``packages`` namespace (packages/dp_tokenize.py,
packages/tokenizer_utils.py): put ``dp-tokenization_b200/`` on ``sys.path`` and the reference's
callers (``tests/test_tokenization_algorithms.py:7-8``, ``main_analyze_s2orc.py:27``,
``main_biomed_translation.py:25``) import the B200 path unchanged.  See INTEGRATION.md."""
