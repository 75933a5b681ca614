"""Minimal stand-in for the ``prettytable`` wheel (absent from this image): the reference's managers only build
tables for their ``__str__``.  Test infrastructure (tests/ref_runner.py puts this directory on the path)."""


class PrettyTable:
  def __init__(self, field_names=None, **_):
    self.title = ""
    self.field_names = list(field_names or [])
    self.align = {}
    self.rows = []

  def add_row(self, row, **_):
    self.rows.append(list(row))

  def add_rows(self, rows):
    for r in rows:
      self.add_row(r)

  def get_string(self, **_):
    head = " | ".join(map(str, self.field_names))
    body = "\n".join(" | ".join(map(str, r)) for r in self.rows)
    return f"{self.title}\n{head}\n{body}"

  __str__ = get_string
