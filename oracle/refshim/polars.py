"""placeholder (test infrastructure)"""
