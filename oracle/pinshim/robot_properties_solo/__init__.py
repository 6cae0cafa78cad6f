"""stand-in package: the reference's motion files import Solo12Config only to size their IK weight vectors"""
