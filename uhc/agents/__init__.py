from uhc.agents.agent_copycat import AgentCopycat

agent_dict = {"agent_copycat": AgentCopycat}
