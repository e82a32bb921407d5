"""reference src/akshar/tokenizer.py -> akshar_b200.tokenizer"""
from akshar_b200.tokenizer import aksharTokenizer, AksharTokenizer  # noqa: F401
