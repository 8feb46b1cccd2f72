"""ET examples as integration tests (reference ExampleTest / TableAccessTest / PlanExecutorTest /
DataLoadTest run the same scenarios on the REEF local runtime)."""
import pytest

from harmony_b200.et import examples


@pytest.mark.integration
@pytest.mark.parametrize("name", sorted(examples.ALL))
def test_example(name):
    assert examples.ALL[name]()
