NO = "no"
