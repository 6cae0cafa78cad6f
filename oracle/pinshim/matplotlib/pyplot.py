"""inert stand-in"""
