"""Test suite for Akshar tokenizer."""

